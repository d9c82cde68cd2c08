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


def tetra_sharded(engine, genome_ids: Sequence[int], genome_lengths: Sequence[int], rank: int, world: int, *, device=None, group=None,
                  compute=None) -> Tuple[np.ndarray, np.ndarray]:
    """TETRA across ranks: every rank counts and scores a contiguous, length-balanced block of the genomes,
    one all-gather of the 256 fp64 Z-scores + 256 "observed" flags per genome gives every rank the full Z matrix; the
    N x N correlation (a few ms for N = 1000) is then computed wherever it is needed with ``engine.tetra_corr``.

    -> (z[n][256], observed[n][256]) in the order of ``genome_ids``.  ``compute`` (tests only) replaces
    ``engine.tetra_zscores``."""
    import torch
    import torch.distributed as dist

    ids = np.ascontiguousarray(genome_ids, dtype=np.int32)
    blocks = partition(np.asarray([genome_lengths[i] for i in ids], dtype=np.float64), world)
    s, e = blocks[rank]
    fn = compute if compute is not None else engine.tetra_zscores
    if e > s:
        z, obs = fn(ids[s:e])
    else:
        z, obs = np.zeros((0, 256), np.float64), np.zeros((0, 256), np.uint8)
    if world == 1:
        return z, obs
    pad = max(b - a for a, b in blocks)
    row = 256 * 8 + 256  # bytes per genome: Z-scores then flags
    buf = np.zeros((pad, row), dtype=np.uint8)
    buf[: e - s, : 256 * 8] = np.ascontiguousarray(z, dtype=np.float64).view(np.uint8).reshape(e - s, 256 * 8)
    buf[: e - s, 256 * 8:] = np.ascontiguousarray(obs, dtype=np.uint8)
    t = torch.from_numpy(buf.reshape(-1))
    if device is not None:
        t = t.to(device)
    out = torch.empty(world * pad * row, dtype=torch.uint8, device=t.device)
    dist.all_gather_into_tensor(out, t, group=group)
    flat = out.cpu().numpy().reshape(world, pad, row)
    zs = [flat[r, : b - a, : 256 * 8].copy().view(np.float64).reshape(b - a, 256) for r, (a, b) in enumerate(blocks)]
    os_ = [flat[r, : b - a, 256 * 8:].copy() for r, (a, b) in enumerate(blocks)]
    return np.concatenate(zs), np.concatenate(os_)
