ncu --set full --clock-control none --import-source on -k regex:ext_dp -s 3 -c 1 -o gpurun_out/prof_ext_v4 -f python bench.py --steps 1 --warmup 1 --no-cpu --ncu > gpurun_out/ncu_ext_v4.log 2>&1
python bench.py --no-cpu --steps 2 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step']); print(d['stage_ms_per_step'])"
