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


# ------------------------------------------------------------------------------------------------------------------
# Large sets (BASELINE.json configs[2..4]): every clade is generated from its own SeedSequence child, so clades can be
# built by independent worker processes (and by different ranks) and still give one reproducible set.
def _clade_rng(seed: int, *key: int):
    return np.random.default_rng(np.random.SeedSequence(entropy=seed, spawn_key=tuple(key)))


def root_codes(seed: int, length: int) -> np.ndarray:
    rng = _clade_rng(seed, 0)
    return random_codes(rng, length, float(rng.uniform(0.35, 0.70)))


def make_clade(seed: int, c: int, per_clade: int, root: np.ndarray, *, target_len: int = 0, between=(0.075, 0.125), within=(0.001, 0.05),
               indel_ratio: float = 0.1, n_rearr: int = 20, specific=(0.02, 0.08), n_rate: float = 1e-4, max_contigs: int = 3) -> List[SynthGenome]:
    """The `per_clade` strains of clade `c`: the root diverged by 7.5-12.5 % block-wise substitutions (+ indels,
    rearrangements) gives the clade ancestor -- a star phylogeny, so any two clades are 15-25 % apart (inter-clade
    ANI ~ 75-85 %) -- and the strains derive from it like in :func:`make_set`.  `target_len` > 0 takes a window of that
    many bases of the root first (mixed-length sets, configs[4])."""
    rng = _clade_rng(seed, 1, c)
    gc = float(rng.uniform(0.35, 0.70))
    base = root
    if target_len and target_len < len(root):
        s = int(rng.integers(0, len(root) - target_len))
        base = root[s : s + target_len]
    elif target_len > len(root):  # longer than the root: the root followed by clade-specific sequence
        base = np.concatenate([root, random_codes(rng, target_len - len(root), gc)])
    div = float(rng.uniform(*between))
    anc = indels(rng, substitute_blocks(rng, base, div), div * indel_ratio * 0.5, gc)
    anc = rearrange(rng, anc, n_rearr)
    out = []
    for s in range(per_clade):
        srng = _clade_rng(seed, 2, c, s)
        sub = float(srng.uniform(*within))
        g = substitute(srng, anc, sub)
        g = indels(srng, g, sub * indel_ratio, gc)
        g = rearrange(srng, g, n_rearr)
        g = insert_specific(srng, g, float(srng.uniform(*specific)), gc)
        out.append(to_genome(srng, f"c{c:03d}s{s:02d}", g, c, int(srng.integers(1, max_contigs + 1)), n_rate))
    return out


def _clade_job(args):
    seed, c, per_clade, root_len, target_len, outdir = args
    import os

    root = root_codes(seed, root_len)
    for s, g in enumerate(make_clade(seed, c, per_clade, root, target_len=target_len)):
        stem = os.path.join(outdir, f"g{c:04d}_{s:02d}")
        np.save(stem + ".off.tmp.npy", g.contig_off)
        os.replace(stem + ".off.tmp.npy", stem + ".off.npy")
        np.save(stem + ".res.tmp.npy", g.residues)
        os.replace(stem + ".res.tmp.npy", stem + ".res.npy")  # the residues file appears last: it marks the genome complete
    return c


def clade_lengths(seed: int, n_clades: int, length, mixed: bool) -> List[int]:
    """Per-clade genome length: constant, or log-uniform in [length[0], length[1]] (configs[4]: 0.5-12 Mb)."""
    if not mixed:
        return [0] * n_clades
    rng = _clade_rng(seed, 3)
    lo, hi = length
    return [int(x) for x in np.exp(rng.uniform(np.log(lo), np.log(hi), size=n_clades))]


def make_set_cached(n_clades: int, per_clade: int, length, seed: int, *, rank: int = 0, world: int = 1, workers: int = 0,
                    cache_dir: str = "/dev/shm/b200ani_synth", timeout_s: float = 1800.0, only_clades=None) -> List[SynthGenome]:
    """n_clades x per_clade genomes, generated once per machine into `cache_dir` (memory-mapped by every caller).
    Rank r of `world` builds the clades c with c % world == r in a process pool and then waits for the other ranks' files.
    `length`: int (every clade ~length) or (lo, hi) for a log-uniform mixed-length set.  Call before CUDA is initialised
    (the pool forks).  `only_clades`: build / load just these clades (the other entries of the returned list are None)."""
    import os
    import time
    from concurrent.futures import ProcessPoolExecutor
    from multiprocessing import get_context

    mixed = isinstance(length, (tuple, list))
    root_len = int(length[1] if mixed else length)
    tag = f"s{seed}_c{n_clades}x{per_clade}_" + (f"L{length[0]}-{length[1]}" if mixed else f"L{length}")
    outdir = os.path.join(cache_dir, tag)
    os.makedirs(outdir, exist_ok=True)
    tl = clade_lengths(seed, n_clades, length, mixed)
    wanted = set(range(n_clades)) if only_clades is None else set(only_clades)
    mine = [c for c in sorted(wanted) if c % world == rank and
            not all(os.path.exists(os.path.join(outdir, f"g{c:04d}_{s:02d}.res.npy")) for s in range(per_clade))]
    if mine:
        nw = workers or max(1, min(len(mine), (os.cpu_count() or 1) // max(world, 1), 32))
        jobs = [(seed, c, per_clade, root_len, tl[c], outdir) for c in mine]
        if nw == 1:
            for j in jobs:
                _clade_job(j)
        else:
            with ProcessPoolExecutor(nw, mp_context=get_context("fork")) as ex:
                list(ex.map(_clade_job, jobs))
    out: List[SynthGenome] = []
    t0 = time.time()
    for c in range(n_clades):
        for s in range(per_clade):
            if c not in wanted:
                out.append(None)
                continue
            stem = os.path.join(outdir, f"g{c:04d}_{s:02d}")
            while not os.path.exists(stem + ".res.npy"):
                if time.time() - t0 > timeout_s:
                    raise TimeoutError(f"synthetic genome {stem} was not produced by its rank")
                time.sleep(0.05)
            out.append(SynthGenome(f"c{c:03d}s{s:02d}", np.load(stem + ".res.npy", mmap_mode="r"), np.load(stem + ".off.npy"), c))
    return out
