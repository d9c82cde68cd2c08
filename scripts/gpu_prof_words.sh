set -x
cd $GRAFT_REPO_ROOT
export B200ANI_PIPELINES=1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:word_probe_kernel -s 1 -c 1 -o gpurun_out/prof_anib_word_probe_kernel -f \
    python bench.py --workload anib --genomes 20 --steps 1 --warmup 1 --no-cpu --ncu > gpurun_out/ncu_words.log 2>&1
tail -3 gpurun_out/ncu_words.log
ls -la gpurun_out/*.ncu-rep
