set -x
cd $GRAFT_REPO_ROOT
timeout 600 python bench.py --impl reference --steps 3 --warmup 0 > gpurun_out/r02_reference_n1.json 2> gpurun_out/r02_reference_n1.err; tail -2 gpurun_out/r02_reference_n1.err; head -c 900 gpurun_out/r02_reference_n1.json
timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -2 gpurun_out/r02_bench_n1.err; head -c 600 gpurun_out/r02_bench_n1.json
timeout 600 python bench.py --workload anim50 --steps 5 --warmup 3 > gpurun_out/r02_bench_anim50.json 2> gpurun_out/r02_bench_anim50.err; tail -2 gpurun_out/r02_bench_anim50.err; head -c 600 gpurun_out/r02_bench_anim50.json
