#!/usr/bin/env python
"""Scale sanity run (not a benchmark of record): many genomes, all three methods, one GPU.

  python scripts/scale_check.py [n_clades per_clade genome_len]
Prints wall-clock per method; checks a random sample of pairs against the oracle."""
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import oracle  # noqa: E402
import pyani_b200  # noqa: E402
from pyani_b200 import synth, tetra  # noqa: E402

n_clades, per_clade, glen = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (12, 10, 2_000_000)
t0 = time.perf_counter()
gs = synth.make_set(n_clades, per_clade, glen, seed=20260928, within=(0.001, 0.08))
print(f"generated {len(gs)} genomes, {sum(len(g.residues) for g in gs)/1e6:.0f} Mb in {time.perf_counter()-t0:.1f}s", flush=True)
eng = pyani_b200.Engine(0)
t0 = time.perf_counter()
gids = [eng.add_genome(g.residues, g.contig_off, g.name) for g in gs]
print(f"staged in {time.perf_counter()-t0:.2f}s")
pairs = synth.all_directed_pairs(len(gs))
for rep in range(2):
    eng.reset_indexes()
    t0 = time.perf_counter()
    res = eng.anim(pairs)
    dt = time.perf_counter() - t0
    ms, cnt = eng.last_profile()
    print(f"ANIm {len(pairs)} pairs: {dt:.2f}s = {len(pairs)/dt:.0f} comparisons/s; aligned {(res['status']==0).sum()}; stages {{k: round(v) for k, v in ms.items()}}".replace("{k: round(v) for k, v in ms.items()}", str({k: round(v) for k, v in ms.items()})), flush=True)
rng = np.random.default_rng(1)
og = {}
for i in rng.choice(len(pairs), size=6, replace=False):
    a, b = pairs[i]
    for x in (a, b):
        if x not in og:
            og[x] = oracle.OGenome(gs[x].residues, gs[x].contig_off)
    want = oracle.anim_pair(og[a], og[b])[0]
    assert all(int(res[i][k]) == v for k, v in want.items()), (a, b)
print("ANIm sample == oracle")
sub = pairs[(pairs[:, 0] < 40) & (pairs[:, 1] < 40)]
t0 = time.perf_counter()
rb = eng.anib(sub)
dt = time.perf_counter() - t0
print(f"ANIb {len(sub)} pairs: {dt:.2f}s = {len(sub)/dt:.0f} comparisons/s; pass fraction {rb['n_pass'].sum()/rb['n_frags'].sum():.2f}", flush=True)
for i in rng.choice(len(sub), size=4, replace=False):
    a, b = sub[i]
    for x in (a, b):
        if x not in og:
            og[x] = oracle.OGenome(gs[x].residues, gs[x].contig_off)
    want = oracle.anib_pair(og[a], og[b])[0]
    assert all(int(rb[i][k]) == v for k, v in want.items()), (a, b)
print("ANIb sample == oracle")
t0 = time.perf_counter()
z, obs = eng.tetra_zscores(gids)
corr = eng.tetra_corr(z)
dt = time.perf_counter() - t0
print(f"TETRA {len(gids)} genomes: Z + {len(gids)*(len(gids)-1)//2} correlations in {dt*1e3:.1f} ms")
c2, c3, c4 = oracle.tetra_counts(gs[3].residues, gs[3].contig_off)
assert (oracle.tetra_zscores(c2, c3, c4)[0] == z[3]).all()
print("TETRA sample == oracle")
