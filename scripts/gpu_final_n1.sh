set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 600 python bench.py --impl reference --steps 3 --warmup 0 > gpurun_out/r02_reference_n1.json 2> gpurun_out/r02_reference_n1.err; tail -2 gpurun_out/r02_reference_n1.err; head -c 400 gpurun_out/r02_reference_n1.json
timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -2 gpurun_out/r02_bench_n1.err; head -c 400 gpurun_out/r02_bench_n1.json
timeout 600 python bench.py --workload anim50 --steps 5 --warmup 3 > gpurun_out/r02_bench_anim50.json 2> gpurun_out/r02_bench_anim50.err; tail -2 gpurun_out/r02_bench_anim50.err; head -c 400 gpurun_out/r02_bench_anim50.json
timeout 900 python bench.py --workload tetra --steps 5 --warmup 3 > gpurun_out/r02_bench_tetra.json 2> gpurun_out/r02_bench_tetra.err; tail -2 gpurun_out/r02_bench_tetra.err; head -c 400 gpurun_out/r02_bench_tetra.json
timeout 1500 python bench.py --workload anib --steps 2 --warmup 3 > gpurun_out/r02_bench_anib.json 2> gpurun_out/r02_bench_anib.err; tail -2 gpurun_out/r02_bench_anib.err; head -c 400 gpurun_out/r02_bench_anib.json
export B200ANI_PIPELINES=1
for k in gap_dp_kernel seed_probe_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o gpurun_out/prof_anim50_$k -f \
      python bench.py --workload anim50 --steps 1 --warmup 1 --no-cpu --ncu > gpurun_out/ncu_$k.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
