#!/bin/bash
# Run under gpurun (1 GPU).  Produces gpurun_out/{launches.csv, prof_*.ncu-rep}; summarise with profiles/summarise.py.
# Numbers printed by a bench run under ncu are never bench values.
set -x
mkdir -p gpurun_out
# (1) launch list: every kernel launch of a short bench run with its device time
ncu --metrics gpu__time_duration.sum --clock-control none -s ${NSKIP:-900} -c ${NLAUNCH:-1500} --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu --ncu > gpurun_out/bench_under_ncu.log 2>&1
# (2) full capture of the DP kernels and the seed kernel (a few launches each)
for k in gap_dp_kernel ext_dp seed_probe seed_extend; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s ${SKIP:-2} -c 2 -o gpurun_out/prof_$k -f \
      python bench.py --steps 1 --warmup 1 --no-cpu --ncu > gpurun_out/ncu_$k.log 2>&1
done
ls -la gpurun_out
