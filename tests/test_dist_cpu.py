"""Sharding logic of pyani_b200/dist.py on CPU: world_size-2 gloo processes, no GPU."""
import os
import socket

import numpy as np
import pytest

from pyani_b200 import dist as pdist
from pyani_b200.engine import ANIB_DTYPE, ANIM_DTYPE


def test_partition_is_contiguous_covering_and_balanced():
    rng = np.random.default_rng(0)
    costs = rng.uniform(1, 10, size=1000)
    for world in (1, 2, 3, 8):
        blocks = pdist.partition(costs, world)
        assert len(blocks) == world and blocks[0][0] == 0 and blocks[-1][1] == 1000
        assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
        sums = [costs[s:e].sum() for s, e in blocks]
        assert max(sums) - min(sums) <= 2 * costs.max()
    assert pdist.partition(np.ones(3), 8)[-1][1] == 3  # more ranks than pairs: empty blocks are fine


def _fake_compute(pairs):
    out = np.zeros(len(pairs), dtype=ANIM_DTYPE)
    out["ref_aln_len"] = pairs[:, 0] * 1000 + pairs[:, 1]
    out["qry_aln_len"] = pairs[:, 1] * 1000 + pairs[:, 0]
    out["sum_aligned"] = 2 * out["ref_aln_len"]
    out["status"] = (pairs[:, 0] + pairs[:, 1]) % 2
    return out


def _fake_compute_b(pairs):
    out = np.zeros(len(pairs), dtype=ANIB_DTYPE)
    out["aln_length"] = pairs[:, 0] * 1000 + pairs[:, 1]
    out["pid_milli_sum"] = 7 * out["aln_length"]
    out["n_frags"] = pairs[:, 0] + 1
    out["n_pass"] = pairs[:, 1]
    return out


def _fake_tetra(ids):
    ids = np.asarray(ids)
    z = (ids[:, None] * 1000.0 + np.arange(256)[None, :]) / 7.0
    obs = ((ids[:, None] + np.arange(256)[None, :]) % 3 != 0).astype(np.uint8)
    return z, obs


def _worker(rank, world, port, q):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pyani_b200 import synth

    pairs = synth.all_directed_pairs(7)
    lengths = [5, 1, 9, 3, 3, 7, 2]
    full = pdist.anim_sharded(None, pairs, lengths, rank, world, compute=_fake_compute)
    fullb = pdist.anib_sharded(None, pairs, lengths, rank, world, compute=_fake_compute_b)
    z, obs = pdist.tetra_sharded(None, list(range(7)), lengths, rank, world, compute=_fake_tetra)
    q.put((rank, full.tobytes() + fullb.tobytes() + z.tobytes() + obs.tobytes()))
    dist.destroy_process_group()


def test_two_rank_gloo_all_gather_assembles_full_result():
    import torch.multiprocessing as mp

    from pyani_b200 import synth

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    zw, ow = _fake_tetra(list(range(7)))
    want = _fake_compute(synth.all_directed_pairs(7)).tobytes() + _fake_compute_b(synth.all_directed_pairs(7)).tobytes() + zw.tobytes() + ow.tobytes()
    assert got[0] == want and got[1] == want  # every rank holds the full, correctly ordered array


def test_deal_partitions_and_keeps_directions_together():
    from pyani_b200 import synth

    pairs = synth.all_directed_pairs(23)
    for world in (1, 2, 3, 8):
        shares = pdist.deal(pairs, world, chunk=4)
        allidx = np.concatenate(shares)
        assert sorted(allidx.tolist()) == list(range(len(pairs)))
        owner = np.empty(len(pairs), int)
        for r, s in enumerate(shares):
            owner[s] = r
        lut = {(a, b): owner[i] for i, (a, b) in enumerate(pairs.tolist())}
        assert all(lut[(a, b)] == lut[(b, a)] for a, b in lut)  # both directions of a genome pair on one rank
        if world > 1:
            sizes = [len(s) for s in shares]
            assert max(sizes) - min(sizes) <= 2 * 4 * 2  # at most one chunk (of 4 unordered pairs) apart, plus the tail
        # walk order inside a share: (lower id, higher id) ascending, the two directions adjacent
        s0 = pairs[shares[0]]
        keys = [(min(a, b), max(a, b)) for a, b in s0.tolist()]
        assert keys == sorted(keys)


def test_default_dealing_balances_clade_structured_costs():
    """Genome ids grouped by clade (10 strains each), a within-clade comparison 20 x as costly as a between-clade one: the
    default deal (single unordered pairs, round-robin in walk order) keeps every rank within 1 % of the mean cost -- chunks
    of 64 hand the clade mates of a genome to one rank as a lump (several per cent apart, what the 8-GPU run measured)."""
    from pyani_b200 import synth

    n, world = 200, 8
    pairs = synth.all_directed_pairs(n)
    cost = np.where(pairs[:, 0] // 10 == pairs[:, 1] // 10, 20.0, 1.0)

    def spread(chunk):
        tot = np.array([cost[s].sum() for s in pdist.deal(pairs, world, chunk=chunk)])
        return tot.max() / tot.mean()

    assert spread(1) == pytest.approx(spread(pdist.deal.__defaults__[0]))  # the default IS single dealing
    assert spread(1) < 1.01 < spread(64)


def test_bench_cpu_sample_keeps_every_core_busy_with_the_full_lists_clade_mix():
    import importlib.util
    from pathlib import Path

    spec = importlib.util.spec_from_file_location("bench_mod", Path(__file__).resolve().parents[1] / "bench.py")
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    idx, pairs, within, full_within = bench.cpu_sample(50, 10, 128)  # ANIm, configs[3] on a 128-thread host
    assert 2 * 128 <= len(pairs) <= 6 * 128 and len(pairs) == len(idx) * (len(idx) - 1)
    assert abs(within / len(pairs) - full_within) < 0.005
    idx, pairs, within, full_within = bench.cpu_sample(20, 10, 128, 102, 153)  # ANIb: one pair per thread
    assert 102 <= len(pairs) <= 153 and abs(within / len(pairs) - full_within) < 0.01


def test_index_layout_blocks_are_disjoint_aligned_and_equal_sized():
    """dist.index_layout (the slab the ranks all-gather their seed indexes through): every genome's four arrays lie inside its
    owner's block, 256-byte aligned, without overlap; table sizes follow the engine's rule (smallest power of two >= 2 x minimizers)."""
    rng = np.random.default_rng(3)
    n_mini = rng.integers(1_000, 2_000_000, size=37)
    for world in (1, 2, 8):
        block_of, off, bits, block_bytes = pdist.index_layout(n_mini, world)
        assert (np.diff(block_of) >= 0).all() and block_of.max() <= world - 1
        assert ((1 << bits) >= 2 * n_mini).all() and ((1 << (bits - 1)) < 2 * n_mini).all()
        sizes = np.stack([4 * n_mini, 4 * n_mini, 8 << bits, 4 << (bits - 2)], axis=1)
        spans = sorted((int(off[g, a]), int(off[g, a] + sizes[g, a]), g) for g in range(len(n_mini)) for a in range(4))
        assert all(s % 256 == 0 for s, _, _ in spans)
        assert all(spans[i][1] <= spans[i + 1][0] for i in range(len(spans) - 1))
        for s, e, g in spans:
            assert block_of[g] * block_bytes <= s and e <= (block_of[g] + 1) * block_bytes
