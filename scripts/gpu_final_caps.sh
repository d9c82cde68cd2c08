set -x
cd $GRAFT_REPO_ROOT
export B200ANI_PIPELINES=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tetra_count_kernel -s 1 -c 1 -o gpurun_out/prof_tetra_count_kernel -f \
    python bench.py --workload tetra --genomes 100 --steps 1 --warmup 1 --no-cpu --ncu > gpurun_out/ncu_tetra_count.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'ext_dp_kernel|seed_probe_kernel|gap_dp_kernel' -s 12 -c 12 -o gpurun_out/prof_anim_kernels -f \
    python bench.py --workload anim --steps 1 --warmup 1 --no-cpu --ncu > gpurun_out/ncu_anim.log 2>&1
ls -la gpurun_out/*.ncu-rep
