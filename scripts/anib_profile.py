import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, pyani_b200
from pyani_b200 import synth
gs = synth.make_set(4, 10, 2_000_000, seed=20260928, within=(0.001, 0.08))
eng = pyani_b200.Engine(0)
gids = [eng.add_genome(g.residues, g.contig_off, g.name) for g in gs]
pairs = synth.all_directed_pairs(len(gs))
for rep in range(2):
    t0 = time.perf_counter(); rb = eng.anib(pairs); dt = time.perf_counter() - t0
    ms, cnt = eng.last_profile()
    print(f"ANIb {len(pairs)} pairs {dt:.2f}s", {k: round(v, 1) for k, v in ms.items() if v}, {k: v for k, v in cnt.items() if v})
