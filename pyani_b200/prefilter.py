"""Opt-in sketch prefilter: skip pairs that cannot align before any alignment work is spent on them.

The reference has no prefilter: ANIm on unrelated genomes ends in ``sys.exit("Zero length alignment!")`` (anim.py:481-486);
its fastANI wrapper (fastani.py:193-228) is the pattern for "estimate first".  Here every staged genome has a bottom-s MinHash
sketch of its seed k-mers (``Engine.sketch``: the s smallest distinct ``fmix32`` hashes of its canonical 16-mers, computed on
the device from the seed index), two sketches give a Mash-style Jaccard / identity estimate on the host, and pairs below a
threshold are left out of the engine call (they would come back with ``status = no_alignment``).  Because it changes which
pairs are computed it is never applied unless asked for."""
from __future__ import annotations

from typing import Dict, Sequence

import numpy as np

K = 16  # seed k-mer length of the ANIm index (AN_K)


def jaccard(sa: np.ndarray, sb: np.ndarray) -> float:
    """Mash's estimator: shared values among the bottom-s of the UNION of the two sketches."""
    if len(sa) == 0 or len(sb) == 0:
        return 0.0
    s = min(len(sa), len(sb))
    union = np.union1d(sa, sb)[:s]
    shared = np.intersect1d(sa, sb, assume_unique=True)
    return float(np.count_nonzero(shared <= union[-1])) / float(len(union))


def identity_estimate(j: float, k: int = K) -> float:
    """Mash distance -> identity: 1 + ln(2j / (1 + j)) / k (0 for disjoint sketches)."""
    if j <= 0.0:
        return 0.0
    return max(0.0, 1.0 + np.log(2.0 * j / (1.0 + j)) / k)


def sketches(engine, gids: Sequence[int], s: int = 2048) -> Dict[int, np.ndarray]:
    return {int(g): engine.sketch(int(g), s) for g in gids}


def pair_mask(engine, pairs: np.ndarray, min_identity: float = 0.70, s: int = 2048) -> np.ndarray:
    """Boolean mask over the rows of ``pairs`` (genome ids): True = the sketches put the pair at or above ``min_identity``.
    Both directions of a pair get the same answer.  ANIm aligns down to ~75 % identity; 0.70 leaves a safety margin."""
    pairs = np.asarray(pairs).reshape(-1, 2)
    sk = sketches(engine, np.unique(pairs), s)
    cache: Dict[tuple, bool] = {}
    out = np.zeros(len(pairs), dtype=bool)
    for i, (a, b) in enumerate(pairs.tolist()):
        key = (a, b) if a <= b else (b, a)
        if key not in cache:
            cache[key] = a == b or identity_estimate(jaccard(sk[a], sk[b])) >= min_identity
        out[i] = cache[key]
    return out
