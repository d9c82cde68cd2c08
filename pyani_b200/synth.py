"""Synthetic, phylogenetically structured genome sets for benchmarks and parity tests.

Unrelated random genomes are useless for ANI (zero alignments; the reference aborts, anim.py:481-486),
so sets are built as clades: a root sequence, clade ancestors diverged from it by substitution, and
strains derived from their clade ancestor by substitutions, small indels, inversions / translocations,
strain-specific insertions and a sprinkle of Ns.  Everything is seeded (NumPy PCG64) and reproducible.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Tuple

import numpy as np

_ASCII = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.array([3, 2, 1, 0], dtype=np.uint8)


@dataclass
class SynthGenome:
    name: str
    residues: np.ndarray  # uint8 ASCII
    contig_off: np.ndarray  # uint64 [n+1]
    clade: int


def random_codes(rng, n: int, gc: float) -> np.ndarray:
    p = np.array([(1 - gc) / 2, gc / 2, gc / 2, (1 - gc) / 2])
    return rng.choice(4, size=n, p=p).astype(np.uint8)


def substitute(rng, codes: np.ndarray, rate: float) -> np.ndarray:
    out = codes.copy()
    if rate <= 0:
        return out
    pos = np.flatnonzero(rng.random(len(codes)) < rate)
    out[pos] = (out[pos] + rng.integers(1, 4, size=len(pos)).astype(np.uint8)) & 3
    return out


def substitute_blocks(rng, codes: np.ndarray, mean_rate: float, block: int = 1000, shape: float = 2.0) -> np.ndarray:
    """Substitutions whose rate varies per `block` bases (gamma distributed, mean `mean_rate`): conserved and
    diverged regions alternate as in real genomes, so distant clades still share alignable stretches."""
    nb = (len(codes) + block - 1) // block
    rates = np.minimum(rng.gamma(shape, mean_rate / shape, size=nb), 0.75)
    per_base = np.repeat(rates, block)[: len(codes)]
    out = codes.copy()
    pos = np.flatnonzero(rng.random(len(codes)) < per_base)
    out[pos] = (out[pos] + rng.integers(1, 4, size=len(pos)).astype(np.uint8)) & 3
    return out


def indels(rng, codes: np.ndarray, rate: float, gc: float, mean_len: float = 3.0) -> np.ndarray:
    n_ev = int(len(codes) * rate)
    if n_ev == 0:
        return codes
    pos = np.sort(rng.integers(0, len(codes), size=n_ev))
    lens = rng.geometric(1.0 / mean_len, size=n_ev)
    is_del = rng.random(n_ev) < 0.5
    keep = np.ones(len(codes), dtype=bool)
    for p, l in zip(pos[is_del], lens[is_del]):
        keep[p : p + l] = False
    ins_pos, ins_vals = [], []
    for p, l in zip(pos[~is_del], lens[~is_del]):
        ins_pos.extend([p] * l)
        ins_vals.append(random_codes(rng, l, gc))
    out = codes
    if ins_pos:
        kept_before = np.cumsum(keep) - keep  # index mapping after deletions
        ip = kept_before[np.asarray(ins_pos)]
        out = np.insert(codes[keep], ip, np.concatenate(ins_vals))
    else:
        out = codes[keep]
    return out


def rearrange(rng, codes: np.ndarray, n_events: int, max_len: int = 50_000) -> np.ndarray:
    for _ in range(n_events):
        L = len(codes)
        ln = int(rng.integers(1000, max(1001, min(max_len, L // 4))))
        s = int(rng.integers(0, L - ln))
        seg = codes[s : s + ln]
        if rng.random() < 0.5:  # inversion in place
            codes = np.concatenate([codes[:s], _COMP[seg[::-1]], codes[s + ln :]])
        else:  # translocation
            rest = np.concatenate([codes[:s], codes[s + ln :]])
            t = int(rng.integers(0, len(rest)))
            codes = np.concatenate([rest[:t], seg, rest[t:]])
    return codes


def insert_specific(rng, codes: np.ndarray, frac: float, gc: float, n_blocks: int = 8) -> np.ndarray:
    total = int(len(codes) * frac)
    if total <= 0:
        return codes
    sizes = rng.multinomial(total, np.ones(n_blocks) / n_blocks)
    for sz in sizes:
        t = int(rng.integers(0, len(codes)))
        codes = np.concatenate([codes[:t], random_codes(rng, int(sz), gc), codes[t:]])
    return codes


def to_genome(rng, name: str, codes: np.ndarray, clade: int, n_contigs: int, n_rate: float) -> SynthGenome:
    res = _ASCII[codes]
    if n_rate > 0:
        res = res.copy()
        res[rng.random(len(res)) < n_rate] = ord("N")
    cuts = np.sort(rng.choice(np.arange(1000, len(res) - 1000), size=n_contigs - 1, replace=False)) if n_contigs > 1 else np.array([], dtype=np.int64)
    off = np.concatenate([[0], cuts, [len(res)]]).astype(np.uint64)
    return SynthGenome(name, np.ascontiguousarray(res), off, clade)


def make_set(n_clades: int, per_clade: int, length: int, seed: int, *, between=(0.15, 0.25), within=(0.001, 0.05),
             indel_ratio: float = 0.1, n_rearr: int = 20, specific=(0.02, 0.08), n_rate: float = 1e-4,
             max_contigs: int = 3, length_jitter: float = 0.1) -> List[SynthGenome]:
    """n_clades x per_clade genomes of ~`length` bases (config 2 of BASELINE.json: 5 x 10 x 5 Mb)."""
    rng = np.random.default_rng(seed)
    out: List[SynthGenome] = []
    for c in range(n_clades):
        gc = float(rng.uniform(0.35, 0.70))
        L = int(length * rng.uniform(1 - length_jitter, 1 + length_jitter))
        anc = random_codes(rng, L, gc)
        if c > 0 and out:
            # clades share a common root: diverge the first clade's ancestor instead of independent noise
            root = _root_cache["codes"]
            div = float(rng.uniform(*between))
            anc = indels(rng, substitute_blocks(rng, root, div), div * indel_ratio * 0.5, gc)
            anc = rearrange(rng, anc, n_rearr)
        else:
            _root_cache["codes"] = anc
        for s in range(per_clade):
            sub = float(rng.uniform(*within))
            g = substitute(rng, anc, sub)
            g = indels(rng, g, sub * indel_ratio, gc)
            g = rearrange(rng, g, n_rearr)
            g = insert_specific(rng, g, float(rng.uniform(*specific)), gc)
            out.append(to_genome(rng, f"c{c:02d}s{s:02d}", g, c, int(rng.integers(1, max_contigs + 1)), n_rate))
    return out


_root_cache = {}


def all_directed_pairs(n: int) -> np.ndarray:
    """itertools.permutations(range(n), 2) as an (n*(n-1), 2) int32 array, grouped by reference
    (subcmd_anim.py:232; both directions are separate comparisons)."""
    a, b = np.meshgrid(np.arange(n, dtype=np.int32), np.arange(n, dtype=np.int32), indexing="ij")
    m = a != b
    return np.stack([a[m], b[m]], axis=1)
