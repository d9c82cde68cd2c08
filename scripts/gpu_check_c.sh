set -x
cd $GRAFT_REPO_ROOT
timeout 900 python scripts/anib_word_debug.py 2>&1 | tail -12
timeout 1500 python -m pytest tests/test_anib_gpu.py tests/test_concordance_jspecies.py -m gpu -q 2>&1 | tail -15
timeout 1200 python bench.py --workload anib --steps 1 --warmup 3 --no-cpu > gpurun_out/r02_anib200_e.json 2> gpurun_out/r02_anib200_e.err; tail -3 gpurun_out/r02_anib200_e.err; head -c 300 gpurun_out/r02_anib200_e.json
