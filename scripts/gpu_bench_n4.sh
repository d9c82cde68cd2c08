export NGPU=4 MIXED=0
bash scripts/gpu_benchN.sh
