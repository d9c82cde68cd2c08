set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,memory.total --format=csv
nproc; free -g | head -2; df -h /dev/shm | tail -1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 1500 python bench.py --workload anim --steps 1 --warmup 3 > gpurun_out/r02_anim500_a.json 2> gpurun_out/r02_anim500_a.err; tail -3 gpurun_out/r02_anim500_a.err; head -c 3000 gpurun_out/r02_anim500_a.json
timeout 600 python bench.py --impl reference --workload anim --steps 2 --warmup 0 > gpurun_out/r02_ref500_a.json 2> gpurun_out/r02_ref500_a.err; tail -3 gpurun_out/r02_ref500_a.err; cat gpurun_out/r02_ref500_a.json
