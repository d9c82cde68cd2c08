set -x
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 900 python bench.py --workload tetra --steps 3 --warmup 3 > gpurun_out/r02_tetra_a.json 2> gpurun_out/r02_tetra_a.err; tail -3 gpurun_out/r02_tetra_a.err; head -c 2500 gpurun_out/r02_tetra_a.json
timeout 600 python bench.py --workload anib --genomes 50 --ncu --steps 1 --warmup 1 > gpurun_out/r02_anib50_a.json 2> gpurun_out/r02_anib50_a.err; tail -3 gpurun_out/r02_anib50_a.err; head -c 2500 gpurun_out/r02_anib50_a.json
