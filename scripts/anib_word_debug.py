"""GPU-side check of the ANIb word stage against the oracle's counters (run on the GPU box): single directions first (no
mirror feeding), then both directions of the same pair in one call."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tests"))
import oracle  # noqa: E402
import pyani_b200  # noqa: E402
from conftest import fixture_genomes  # noqa: E402

files = fixture_genomes()[:2]
eng = pyani_b200.Engine(0)
gids, og = [], []
for f in files:
    arr, off, _ = oracle.read_fasta(f)
    gids.append(eng.add_genome(arr, off, f.stem))
    og.append(oracle.OGenome(arr, off))
F = ["aln_length", "sim_errs", "pid_milli_sum", "n_frags", "n_pass", "n_hsps"]
for label, gp in (("0->1", [[0, 1]]), ("1->0", [[1, 0]]), ("both", [[0, 1], [1, 0]])):
    res = eng.anib(np.array(gp, np.int32))
    _, cnt = eng.last_profile()
    print(label, "gpu", [tuple(int(r[f]) for f in F) for r in res], {k: cnt[k] for k in ("word_hits", "word_anchors", "anchors", "segments")}, flush=True)
for q, s in ((0, 1), (1, 0)):
    want, _, st = oracle.anib_pair(og[q], og[s])
    print(f"{q}->{s}", "orc", tuple(want[f] for f in F), {k: st[k] for k in ("word_hits", "word_anchors", "anchors", "segments")}, flush=True)
