"""Multi-GPU sharding of the directed-pair list: one process per GPU, no data-path collective, one
all-gather of the fixed-size per-pair result structs at the end.

The reference distributes the same independent units as one OS process per pair
(run_multiprocessing.py:130-144) or SGE array tasks (run_sge.py:87-112) and exchanges results through
files; here every rank holds all packed genomes (tiny next to 180 GB of HBM), is dealt its share of the
pair list (:func:`deal`: unordered genome pairs round-robin in walk order) and contributes its 48-byte
result structs to a single ``all_gather`` so that every rank ends up with the full matrices.  Optionally
(:func:`share_indexes`) the seed indexes are built once across the ranks -- each rank builds those of 1/N
of the genomes and one all-gather of the index slab hands every rank all of them.
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


def deal(pairs: np.ndarray, world: int, chunk: int = 1) -> List[np.ndarray]:
    """Which rows of ``pairs`` each rank computes: the directed pairs are grouped by unordered genome pair (both directions
    share one seed pass on the device, so they stay together), the groups sorted by (lower id, higher id) -- the order in
    which the engine walks them, one reference table hot in L2 after the other -- are cut into chunks of ``chunk`` groups,
    and the chunks are dealt round-robin.  Costly and cheap comparisons (within- and between-clade pairs differ several-fold
    in alignment work, and genome ids are normally grouped by clade) are thereby spread evenly without a cost model.  The
    default deals single genome pairs: the costly pairs of one genome are its <= ~10 clade mates, which sit next to each
    other in walk order, so any chunk longer than that hands them to one rank as a lump (8 ranks, 500 genomes in clades of
    10: 3.5 % slowest-rank excess measured with chunks of 64, 0.2 % expected with 1); a rank still meets every reference's
    partners consecutively (runs of ~n/2 partners per genome, every world-th one is its own), so the table stays hot.
    Every rank can compute every rank's share, so nothing but the results is ever exchanged.
    -> list of index arrays (ascending walk order), one per rank, together a partition of range(len(pairs))."""
    p = np.asarray(pairs).reshape(-1, 2)
    lo, hi = p.min(axis=1).astype(np.int64), p.max(axis=1).astype(np.int64)
    order = np.lexsort((p[:, 0], hi, lo))
    if world <= 1:
        return [order]
    key = lo[order] * (int(hi.max()) + 1 if len(p) else 1) + hi[order]
    group = np.concatenate([[0], np.cumsum(key[1:] != key[:-1])]) if len(p) else np.zeros(0, np.int64)
    owner = (group // max(1, chunk)) % world
    return [order[owner == r] for r in range(world)]


def all_gather_results(local: np.ndarray, blocks, device=None, group=None, dtype=ANIM_DTYPE) -> np.ndarray:
    """ONE all-gather of the per-rank result arrays (padded to the largest share) -> full array in caller order, on every
    rank.  ``blocks``: per rank either a contiguous (start, end) range of the pair list or the index array of :func:`deal`.
    ``dtype``: the result struct (ANIM_DTYPE 48 B, ANIB_DTYPE 40 B).

    Works with the ``nccl`` backend (tensors on ``device``) and with ``gloo`` (CPU tensors, used by the
    world_size-2 CPU tests)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    assert len(blocks) == world
    idx = [np.arange(b[0], b[1]) if isinstance(b, tuple) else np.asarray(b) for b in blocks]
    pad = max(len(i) for i in idx)
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
    full = np.zeros(sum(len(i) for i in idx), dtype=dtype)
    for r, i in enumerate(idx):
        full[i] = flat[r, : len(i) * item].copy().view(dtype)
    return full


def _sharded(run, pairs: np.ndarray, rank: int, world: int, dtype, device, group, chunk: int) -> np.ndarray:
    pairs = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
    shares = deal(pairs, world, chunk)
    mine = shares[rank]
    local = run(pairs[mine]) if len(mine) else np.zeros(0, dtype=dtype)
    if world == 1:
        full = np.zeros(len(pairs), dtype=dtype)
        full[mine] = local
        return full
    return all_gather_results(local, shares, device=device, group=group, dtype=dtype)


def anim_sharded(engine, pairs: np.ndarray, genome_lengths: Sequence[int], rank: int, world: int, *, maxmatch=False,
                 nofilter=False, device=None, group=None, compute=None, chunk: int = 1) -> np.ndarray:
    """Run this rank's share of ``pairs`` (see :func:`deal`) on ``engine`` -- every rank holds all genomes -- and
    all-gather the full result array in the order of ``pairs``.  ``genome_lengths`` is unused by the dealing plan and kept
    for callers that balance contiguous blocks with :func:`partition` instead.

    ``compute`` (tests only) replaces ``engine.anim`` so the sharding logic can be exercised without a GPU."""
    fn = compute if compute is not None else (lambda p: engine.anim(p, maxmatch=maxmatch, nofilter=nofilter))
    return _sharded(fn, pairs, rank, world, ANIM_DTYPE, device, group, chunk)


def anib_sharded(engine, pairs: np.ndarray, genome_lengths: Sequence[int], rank: int, world: int, *, fragsize: int = 1020,
                 device=None, group=None, compute=None, chunk: int = 1) -> np.ndarray:
    """ANIb counterpart of :func:`anim_sharded`: pairs are (query, subject), the gathered structs are the 40-byte ANIb
    results."""
    fn = compute if compute is not None else (lambda p: engine.anib(p, fragsize=fragsize))
    return _sharded(fn, pairs, rank, world, ANIB_DTYPE, device, group, chunk)


def index_layout(n_mini: np.ndarray, world: int):
    """Where every genome's four index arrays sit in the shared slab: genome g belongs to rank ``block_of[g]`` (contiguous,
    equal-count blocks), every rank's block has the same size (the largest one, what a single all-gather needs).
    -> (block_of, offsets (n, 4) relative to the slab start: kmer, posflag, table, bitmap; table_bits (n,); block_bytes)."""
    n = len(n_mini)
    nm = np.asarray(n_mini, dtype=np.int64)
    bits = np.maximum(4, np.ceil(np.log2(np.maximum(2 * nm, 1))).astype(np.int64))
    bits = np.where((1 << bits) < 2 * nm, bits + 1, bits)  # the engine's rule: smallest power of two >= 2 * n_mini (at least 2^4)
    per = (n + world - 1) // world
    block_of = np.minimum(np.arange(n) // max(per, 1), world - 1)

    def al(x):
        return (x + 255) & ~255

    sizes = np.stack([al(4 * nm), al(4 * nm), al(8 << bits), al(4 << (bits - 2))], axis=1)
    rel = np.zeros((n, 4), dtype=np.int64)
    block_bytes = 0
    for r in range(world):
        o = 0
        for g in np.flatnonzero(block_of == r):
            for a in range(4):
                rel[g, a] = o
                o += int(sizes[g, a])
        block_bytes = max(block_bytes, o)
    offsets = rel + (block_of * block_bytes)[:, None]
    return block_of, offsets, bits, int(block_bytes)


def share_indexes(engine, gids: Sequence[int], rank: int, world: int, device, group=None):
    """Seed indexes of ``gids`` (the same staged genomes on every rank) built ONCE across the ranks: rank r builds its block of
    the genomes, the blocks travel in one all-gather over NVLink, every rank adopts the arrays of the genomes it did not build
    (SURVEY.md 8e).  Replaces rebuilding all indexes on every rank -- the one cost of a sharded run that does not shrink with
    the number of GPUs.  Returns the slab tensor; the caller keeps it alive while the engine uses the adopted indexes (until
    ``reset_indexes`` / ``clear``)."""
    import torch
    import torch.distributed as dist

    gids = [int(g) for g in gids]
    n = len(gids)
    per = (n + world - 1) // world
    mine = [k for k in range(n) if min(k // max(per, 1), world - 1) == rank]
    engine.index_build([gids[k] for k in mine])
    nm_local = torch.zeros(n, dtype=torch.int64, device=device)
    for k in mine:
        nm_local[k] = engine.index_info(gids[k])[0]
    if world > 1:
        dist.all_reduce(nm_local, op=dist.ReduceOp.SUM, group=group)  # disjoint supports: the sum is the union
    n_mini = nm_local.cpu().numpy()
    block_of, off, bits, block_bytes = index_layout(n_mini, world)
    slab = torch.empty(world * block_bytes, dtype=torch.uint8, device=device)
    base = slab.data_ptr()
    for k in mine:
        engine.index_export(gids[k], *(base + int(o) for o in off[k]))
    engine.sync()
    if world > 1:
        torch.cuda.current_stream(device).synchronize()
        dist.all_gather_into_tensor(slab, slab[rank * block_bytes:(rank + 1) * block_bytes].clone(), group=group)
        torch.cuda.current_stream(device).synchronize()
    for k in range(n):
        if block_of[k] != rank:
            engine.index_adopt(gids[k], *(base + int(o) for o in off[k]), int(n_mini[k]), int(bits[k]))
    return slab


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
