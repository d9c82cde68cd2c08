export NGPU=8 MIXED=1
bash scripts/gpu_benchN.sh
