#!/bin/bash
# Round 2 captures.  Run under gpurun (1 GPU): produces gpurun_out/launches.csv and gpurun_out/prof_<workload>_<kernel>.ncu-rep;
# summarise here with `python profiles/summarise.py r02`.  Numbers printed by a bench run under ncu are never bench values.
set -x
cd ${GRAFT_REPO_ROOT:-.}
export B200ANI_PIPELINES=1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_anib_gpu.py -m gpu -x -q 2>&1 | tail -3
# (1) launch list of the headline path on configs[1] (50 genomes: one step = 2 batches) -- every launch with its device time
ncu --metrics gpu__time_duration.sum --clock-control none -s ${NSKIP:-120} -c ${NLAUNCH:-400} --csv --log-file gpurun_out/launches.csv \
    python bench.py --workload anim50 --steps 1 --warmup 1 --no-cpu --ncu > gpurun_out/bench_under_ncu.log 2>&1
# (2) full captures, 2 launches each
for k in ext_dp_kernel gap_dp_kernel seed_probe_kernel; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s ${SKIP:-2} -c 2 -o gpurun_out/prof_anim50_$k -f \
      python bench.py --workload anim50 --steps 1 --warmup 1 --no-cpu --ncu > gpurun_out/ncu_$k.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'tetra_count_kernel|pack_kernel' -s 2 -c 4 -o gpurun_out/prof_tetra_kernels -f \
    python bench.py --workload tetra --genomes 100 --steps 1 --warmup 1 --no-cpu --ncu > gpurun_out/ncu_tetra.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'ext_dp_kernel|word_probe_kernel' -s 2 -c 3 -o gpurun_out/prof_anib_kernels -f \
    python bench.py --workload anib --genomes 20 --steps 1 --warmup 1 --no-cpu --ncu > gpurun_out/ncu_anib.log 2>&1
ls -la gpurun_out/*.ncu-rep
