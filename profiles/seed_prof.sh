for k in seed_probe seed_extend; do
ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o gpurun_out/prof_${k}_v12 -f python bench.py --steps 1 --warmup 1 --no-cpu --ncu > gpurun_out/ncu_${k}_v12.log 2>&1
done
python bench.py --no-cpu --steps 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value']); print(d['stage_ms_per_step'])"
