"""Multi-GPU sharding of the directed-pair list: one process per GPU, no data-path collective, one
all-gather of the fixed-size per-pair result structs at the end.

The reference distributes the same independent units as one OS process per pair
(run_multiprocessing.py:130-144) or SGE array tasks (run_sge.py:87-112) and exchanges results through
files; here every rank holds all packed genomes (tiny next to 180 GB of HBM), takes a contiguous,
cost-balanced block of the pair list and contributes its 48-byte result structs to a single
``all_gather`` so that every rank ends up with the full matrices.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np

from .engine import ANIB_DTYPE, ANIM_DTYPE


def pair_costs(pairs: np.ndarray, genome_lengths: Sequence[int]) -> np.ndarray:
    """Relative cost model of one directed comparison: the query's minimizers are streamed through the
    reference's table and alignment work grows with the shorter genome -> len(ref) + len(qry)."""
    L = np.asarray(genome_lengths, dtype=np.float64)
    p = np.asarray(pairs).reshape(-1, 2)
    return L[p[:, 0]] + L[p[:, 1]]


def partition(costs: np.ndarray, world: int) -> List[Tuple[int, int]]:
    """Contiguous blocks [start, end) of ~equal total cost, one per rank (blocks may be empty)."""
    n = len(costs)
    if world <= 1:
        return [(0, n)]
    cum = np.concatenate([[0.0], np.cumsum(costs, dtype=np.float64)])
    total = cum[-1]
    bounds = [0]
    for r in range(1, world):
        bounds.append(int(np.searchsorted(cum, total * r / world, side="left")))
    bounds.append(n)
    bounds = np.maximum.accumulate(np.minimum(bounds, n))
    return [(int(bounds[r]), int(bounds[r + 1])) for r in range(world)]


def all_gather_results(local: np.ndarray, blocks: List[Tuple[int, int]], device=None, group=None, dtype=ANIM_DTYPE) -> np.ndarray:
    """One all-gather of the per-rank result arrays (padded to the largest block) -> full array, every rank.
    ``dtype``: the result struct (ANIM_DTYPE 48 B, ANIB_DTYPE 40 B).

    Works with the ``nccl`` backend (tensors on ``device``) and with ``gloo`` (CPU tensors, used by the
    world_size-2 CPU tests)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    assert len(blocks) == world
    pad = max(e - s for s, e in blocks)
    item = dtype.itemsize
    buf = np.zeros(pad * item, dtype=np.uint8)
    raw = np.ascontiguousarray(local).view(np.uint8).reshape(-1)
    buf[: len(raw)] = raw
    t = torch.from_numpy(buf)
    if device is not None:
        t = t.to(device)
    out = torch.empty(world * pad * item, dtype=torch.uint8, device=t.device)
    dist.all_gather_into_tensor(out, t, group=group)
    flat = out.cpu().numpy().reshape(world, pad * item)
    parts = [flat[r, : (e - s) * item].copy().view(dtype) for r, (s, e) in enumerate(blocks)]
    return np.concatenate(parts) if parts else np.zeros(0, dtype=dtype)


def anim_sharded(engine, pairs: np.ndarray, genome_lengths: Sequence[int], rank: int, world: int, *, maxmatch=False,
                 nofilter=False, device=None, group=None, compute=None) -> np.ndarray:
    """Run this rank's block of ``pairs`` on ``engine`` and all-gather the full result array.

    ``compute`` (tests only) replaces ``engine.anim`` so the sharding logic can be exercised without a GPU."""
    pairs = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
    blocks = partition(pair_costs(pairs, genome_lengths), world)
    s, e = blocks[rank]
    fn = compute if compute is not None else (lambda p: engine.anim(p, maxmatch=maxmatch, nofilter=nofilter))
    local = fn(pairs[s:e]) if e > s else np.zeros(0, dtype=ANIM_DTYPE)
    if world == 1:
        return local
    return all_gather_results(local, blocks, device=device, group=group)


def anib_sharded(engine, pairs: np.ndarray, genome_lengths: Sequence[int], rank: int, world: int, *, fragsize: int = 1020,
                 device=None, group=None, compute=None) -> np.ndarray:
    """ANIb counterpart of :func:`anim_sharded`: pairs are (query, subject); the cost model is the same (the query's
    fragments stream through the subject's table), the gathered structs are the 40-byte ANIb results."""
    pairs = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
    blocks = partition(pair_costs(pairs, genome_lengths), world)
    s, e = blocks[rank]
    fn = compute if compute is not None else (lambda p: engine.anib(p, fragsize=fragsize))
    local = fn(pairs[s:e]) if e > s else np.zeros(0, dtype=ANIB_DTYPE)
    if world == 1:
        return local
    return all_gather_results(local, blocks, device=device, group=group, dtype=ANIB_DTYPE)
