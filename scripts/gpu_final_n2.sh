export NGPU=2 MIXED=0
bash scripts/gpu_benchN.sh
