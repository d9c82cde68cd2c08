"""Sketch prefilter (opt-in): estimator on CPU, device sketches vs the oracle's minimizers on the GPU."""
import numpy as np
import pytest

from pyani_b200 import prefilter


def _fmix32(h):
    h = h.astype(np.uint64)
    h ^= h >> np.uint64(16); h = (h * np.uint64(0x85ebca6b)) & np.uint64(0xffffffff)
    h ^= h >> np.uint64(13); h = (h * np.uint64(0xc2b2ae35)) & np.uint64(0xffffffff)
    h ^= h >> np.uint64(16)
    return h.astype(np.uint32)


def test_estimator_on_synthetic_sketches():
    rng = np.random.default_rng(3)
    a = np.unique(rng.integers(0, 2**32, 40000, dtype=np.uint64).astype(np.uint32))
    keep = rng.random(len(a)) < 0.5  # half of the k-mers shared
    b = np.unique(np.concatenate([a[keep], rng.integers(0, 2**32, int((~keep).sum()), dtype=np.uint64).astype(np.uint32)]))
    sa, sb = a[:2048], b[:2048]
    j = prefilter.jaccard(sa, sb)
    assert abs(j - 1 / 3) < 0.05          # |A and B| / |A or B| = 0.5 / 1.5
    assert prefilter.jaccard(sa, sa) == 1.0 and prefilter.identity_estimate(1.0) == 1.0
    assert prefilter.jaccard(sa, np.array([], np.uint32)) == 0.0 and prefilter.identity_estimate(0.0) == 0.0
    assert 0.93 < prefilter.identity_estimate(j) < 0.97  # 1 + ln(0.5) / 16


@pytest.mark.gpu
def test_device_sketch_equals_oracle_minimizers_and_masks_unrelated_pairs(engine):
    import oracle
    from pyani_b200 import synth

    gs = synth.make_set(2, 2, 300_000, seed=5)
    rng = np.random.default_rng(9)
    alien = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, 300_000)]
    base = len(engine)
    gids = [engine.add_genome(g.residues, g.contig_off, g.name) for g in gs]
    gids.append(engine.add_genome(alien, np.array([0, len(alien)], np.uint64), "alien"))
    assert gids[0] == base
    for g, gid in zip(gs, gids):
        k, _ = oracle.OGenome(g.residues, g.contig_off).minimizers()
        want = np.unique(_fmix32(k))[:512]
        got = engine.sketch(gid, 512)
        assert got.dtype == np.uint32 and (got == want).all()
    pairs = np.array([[gids[0], gids[1]], [gids[1], gids[0]], [gids[0], gids[4]], [gids[4], gids[2]], [gids[3], gids[3]]], np.int32)
    mask = prefilter.pair_mask(engine, pairs, min_identity=0.70)
    assert mask.tolist() == [True, True, False, False, True]
    res = engine.anim(pairs[[2]])
    assert int(res[0]["status"]) == 1  # what the prefilter spares: a comparison without any alignment
