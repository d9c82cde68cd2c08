export NGPU=8 MIXED=0
bash scripts/gpu_benchN.sh
