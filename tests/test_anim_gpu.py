"""CUDA ANIm path (through the C ABI) vs the CPU oracle: bit-exact on every integer the engine returns."""
import json

import numpy as np
import pytest

import oracle
from conftest import GOLDEN, fixture_genomes

pytestmark = pytest.mark.gpu
FIELDS = ["ref_aln_len", "qry_aln_len", "sum_aligned", "sum_weighted", "sim_errs", "n_alignments", "status"]


@pytest.fixture(scope="module")
def small_set(engine):
    from pyani_b200 import synth

    gs = synth.make_set(3, 2, 300_000, seed=20260924, n_rate=3e-4)
    gids = [engine.add_genome(g.residues, g.contig_off, g.name) for g in gs]
    og = [oracle.OGenome(g.residues, g.contig_off) for g in gs]
    return gs, gids, og


def _check(engine, gids, og, pairs, nofilter):
    gp = np.array([[gids[r], gids[q]] for r, q in pairs], dtype=np.int32)
    res = engine.anim(gp, nofilter=nofilter)
    tot = {}
    for (r, q), got in zip(pairs, res):
        want, _, st = oracle.anim_pair(og[r], og[q], nofilter=nofilter)
        for k, v in st.items():
            tot[k] = tot.get(k, 0) + v
        for f in FIELDS:
            assert int(got[f]) == want[f], (r, q, f, int(got[f]), want[f])
    return res, tot


def test_all_pairs_bit_exact(engine, small_set):
    from pyani_b200 import synth

    gs, gids, og = small_set
    pairs = [tuple(p) for p in synth.all_directed_pairs(len(gs))]
    res, tot = _check(engine, gids, og, pairs, nofilter=False)
    ms, cnt = engine.last_profile()
    # deterministic work counters agree with the oracle's (cells are counted identically)
    for a, b in [("anchors", "anchors"), ("links", "links"), ("cut_links", "cut_links"), ("segments", "segments"),
                 ("gap_cells", "gap_cells"), ("ext_cells", "ext_cells"), ("ext_tiles", "ext_tiles")]:
        assert cnt[a] == tot[b], (a, cnt[a], tot[b])
    assert cnt["launches"] > 0 and cnt["pairs"] == len(pairs)
    assert (res["status"] == 0).sum() >= 6  # within-clade pairs always align
    _check(engine, gids, og, pairs[:8], nofilter=True)


def test_alignment_rows_match(engine, small_set):
    gs, gids, og = small_set
    for r, q in [(0, 1), (1, 0), (0, 2), (4, 1)]:
        for nofilter in (False, True):
            want = oracle.anim_pair(og[r], og[q], nofilter=nofilter)[1]
            got = engine.anim_alignments(gids[r], gids[q], nofilter=nofilter)
            assert got.shape == want.shape and (got == want).all(), (r, q, nofilter)


def test_self_and_unrelated(engine, small_set):
    gs, gids, og = small_set
    rng = np.random.default_rng(3)
    arr = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=150_000)
    off = np.array([0, 150_000], dtype=np.uint64)
    gid = engine.add_genome(arr, off, "unrelated")
    res = engine.anim(np.array([[gids[0], gids[0]], [gids[0], gid], [gid, gids[0]]], dtype=np.int32))
    n0 = int(gs[0].contig_off[-1])
    # a genome against itself: every base aligned once, no errors except around Ns
    assert res[0]["status"] == 0 and res[0]["sim_errs"] == oracle.anim_pair(og[0], og[0])[0]["sim_errs"]
    assert res[0]["ref_aln_len"] == res[0]["qry_aln_len"] <= n0
    assert res[0]["ref_aln_len"] > 0.99 * n0
    # unrelated sequence: soft status instead of the reference's sys.exit (anim.py:481-486)
    assert res[1]["status"] == 1 and res[1]["n_alignments"] == 0 and res[2]["status"] == 1


def test_size_independent_properties(engine):
    """A larger case the oracle is not run on: inserting random records into a genome must not change
    its comparison against the original records (alignments never cross record separators), and
    A->B vs B->A swap the aligned lengths' roles for a pure-substitution pair."""
    from pyani_b200 import synth

    rng = np.random.default_rng(11)
    base = synth.random_codes(rng, 1_500_000, 0.5)
    mut = synth.substitute(rng, base, 0.03)
    a = np.frombuffer(b"ACGT", np.uint8)[base]
    b = np.frombuffer(b"ACGT", np.uint8)[mut]
    junk = np.frombuffer(b"ACGT", np.uint8)[synth.random_codes(rng, 200_000, 0.5)]
    ga = engine.add_genome(a, np.array([0, len(a)], np.uint64))
    gb = engine.add_genome(b, np.array([0, len(b)], np.uint64))
    gb2 = engine.add_genome(np.concatenate([junk, b]), np.array([0, len(junk), len(junk) + len(b)], np.uint64))
    r = engine.anim(np.array([[ga, gb], [gb, ga], [ga, gb2]], np.int32))
    assert r[0]["ref_aln_len"] == r[1]["qry_aln_len"] and r[0]["qry_aln_len"] == r[1]["ref_aln_len"]
    assert r[0]["sim_errs"] == r[1]["sim_errs"]
    for f in FIELDS:
        assert r[0][f] == r[2][f], f
    ident = r[0]["sum_weighted"] / r[0]["sum_aligned"]
    assert abs(ident - 0.97) < 0.002 and r[0]["ref_aln_len"] > 0.995 * len(a)


@pytest.mark.skipif(not fixture_genomes(), reason="fixture genomes not shipped")
def test_reference_fixture_genomes(engine):
    """The four Caulobacter genomes of the reference's ANIm tests: GPU == oracle bit-exact, and both within
    the reference's concordance tolerance of the checked-in nucmer results (see test_anim_oracle_fixtures)."""
    gold = json.load(open(GOLDEN / "anim_parse_delta.json"))
    files = fixture_genomes()
    gids, og = [], []
    for f in files:
        arr, off, _ = oracle.read_fasta(f)
        gids.append(engine.add_genome(arr, off, f.stem))
        og.append(oracle.OGenome(arr, off))
    pairs = [(i, j) for i in range(4) for j in range(4) if i != j]
    res, _ = _check(engine, gids, og, pairs, nofilter=False)
    for (i, j), r in zip(pairs, res):
        g = gold[f"{files[i].stem}_vs_{files[j].stem}.filter"]
        ident = r["sum_weighted"] / r["sum_aligned"]
        assert abs(100 * (ident - g[2])) < 0.1  # the reference's own ANIm tolerance (tests/test_concordance.py:157-159)
        assert abs(r["ref_aln_len"] / g[0] - 1) < 0.015


def test_degenerate_inputs(engine):
    """Empty / tiny / all-N / zero-length-record genomes and an empty pair list must not crash (the
    reference's edge cases are empty and ragged inputs); they yield NO_ALIGNMENT, like the oracle."""
    u8 = lambda s: np.frombuffer(s, dtype=np.uint8)
    rng = np.random.default_rng(5)
    base = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, 5000)]
    cases = {
        "tiny": (u8(b"ACGTACGTAC"), [0, 10]),
        "alln": (u8(b"N" * 3000), [0, 3000]),
        "ragged": (np.concatenate([base[:2000], base[2000:]]), [0, 0, 2000, 2000, 5000]),  # zero-length records
        "normal": (base, [0, 5000]),
        "short_records": (base[:90], [0, 30, 60, 90]),
    }
    gids, og = {}, {}
    for k, (arr, off) in cases.items():
        off = np.asarray(off, dtype=np.uint64)
        gids[k] = engine.add_genome(arr, off, k)
        og[k] = oracle.OGenome(arr, off)
    names = list(cases)
    pairs = [(a, b) for a in names for b in names]
    res = engine.anim(np.array([[gids[a], gids[b]] for a, b in pairs], dtype=np.int32))
    for (a, b), r in zip(pairs, res):
        want = oracle.anim_pair(og[a], og[b])[0]
        for f in FIELDS:
            assert int(r[f]) == want[f], (a, b, f)
    assert res[pairs.index(("normal", "ragged"))]["ref_aln_len"] == 5000  # same bases, different record layout
    assert engine.anim(np.zeros((0, 2), np.int32)).shape == (0,)
    rb = engine.anib(np.array([[gids[a], gids[b]] for a, b in pairs], dtype=np.int32))
    for (a, b), r in zip(pairs, rb):
        want = oracle.anib_pair(og[a], og[b])[0]
        for f in ["aln_length", "sim_errs", "pid_milli_sum", "n_frags", "n_pass", "n_hsps", "status"]:
            assert int(r[f]) == want[f], (a, b, f)
    z, obs = engine.tetra_zscores([gids["tiny"], gids["alln"]])
    assert obs[1].sum() == 0


def test_batch_splitting_on_buffer_overflow():
    """When a batch overflows the anchor / hit buffers the engine retries with half the pairs.  Force that with tiny
    buffers in a fresh process (the capacities are read when the context's workspace is created)."""
    import os
    import subprocess
    import sys

    from conftest import ROOT

    code = r"""
import numpy as np, sys
sys.path.insert(0, %r)
import oracle, pyani_b200
from pyani_b200 import synth
gs = synth.make_set(2, 3, 150_000, seed=77)
with pyani_b200.Engine(0) as eng:
    gids = [eng.add_genome(g.residues, g.contig_off, g.name) for g in gs]
    og = [oracle.OGenome(g.residues, g.contig_off) for g in gs]
    pairs = synth.all_directed_pairs(len(gs))
    res = eng.anim(np.array([[gids[a], gids[b]] for a, b in pairs], dtype=np.int32))
    for (a, b), r in zip(pairs, res):
        want = oracle.anim_pair(og[a], og[b])[0]
        assert all(int(r[k]) == v for k, v in want.items()), (a, b)
    rb = eng.anib(np.array([[gids[a], gids[b]] for a, b in pairs[:8]], dtype=np.int32))
    for (a, b), r in zip(pairs[:8], rb):
        want = oracle.anib_pair(og[a], og[b])[0]
        assert all(int(r[k]) == v for k, v in want.items()), (a, b)
print("SPLIT-OK", int(res["n_alignments"].sum()))
""" % str(ROOT)
    env = dict(os.environ, B200ANI_ANCHOR_CAP="9000", B200ANI_HIT_CAP="20000")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert "SPLIT-OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def test_maxmatch_with_repeats(engine):
    """--maxmatch (anim.py:276-279): seeds may occur several times per genome.  Genomes with 10 copies of a 1.5 kb
    element; CUDA == oracle bit-exact in both modes, and the two modes differ (the repeats only anchor with --maxmatch)."""
    from pyani_b200 import synth

    rng = np.random.default_rng(99)
    elem = synth.random_codes(rng, 1500, 0.5)
    def with_repeats(core, n_copies):
        g = core
        for _ in range(n_copies):
            t = int(rng.integers(0, len(g)))
            g = np.concatenate([g[:t], synth.substitute(rng, elem, 0.01), g[t:]])
        return g
    core = synth.random_codes(rng, 250_000, 0.5)
    a = with_repeats(core, 10)
    b = with_repeats(synth.indels(rng, synth.substitute(rng, core, 0.04), 0.004, 0.5), 10)
    asc = np.frombuffer(b"ACGT", np.uint8)
    arrs = [(asc[a], np.array([0, len(a)], np.uint64)), (asc[b], np.array([0, 120_000, len(b)], np.uint64))]
    gids = [engine.add_genome(x, off, f"rep{i}") for i, (x, off) in enumerate(arrs)]
    pairs = [(0, 1), (1, 0)]
    out = {}
    og = [oracle.OGenome(x, off) for x, off in arrs]
    for mm in (False, True):
        for nofilter in (False, True):
            res = engine.anim(np.array([[gids[r], gids[q]] for r, q in pairs], np.int32), maxmatch=mm, nofilter=nofilter)
            for (r, q), got in zip(pairs, res):
                want = oracle.anim_pair(og[r], og[q], nofilter=nofilter, maxmatch=mm)[0]
                for f in FIELDS:
                    assert int(got[f]) == want[f], (mm, nofilter, r, q, f, int(got[f]), want[f])
            out[(mm, nofilter)] = res
        rows = engine.anim_alignments(gids[0], gids[1], maxmatch=mm, nofilter=True)
        assert (rows == oracle.anim_pair(og[0], og[1], nofilter=True, maxmatch=mm)[1]).all()
    assert out[(True, True)][0]["n_alignments"] > out[(False, True)][0]["n_alignments"]  # repeat copies align to each other
    # ANIb shares the index (seeds unique in both genomes) and still matches the oracle
    rb = engine.anib(np.array([[gids[0], gids[1]]], np.int32))[0]
    want = oracle.anib_pair(og[0], og[1])[0]
    assert all(int(rb[k]) == v for k, v in want.items())


def test_sharded_world1_equals_unsharded_and_shares_compose(engine, small_set):
    """dist.anim_sharded at world 1 == Engine.anim on the same list; and the shares dist.deal hands to 4 ranks, each run on
    the real engine and scattered back by index, give the same array (what the all-gather assembles at N > 1)."""
    from pyani_b200 import dist as pdist
    from pyani_b200 import synth

    gs, gids, og = small_set
    pairs = synth.all_directed_pairs(len(gs))
    gp = np.array([[gids[r], gids[q]] for r, q in pairs], dtype=np.int32)
    direct = engine.anim(gp)
    lengths = [len(g.residues) for g in gs]
    sharded = pdist.anim_sharded(engine, gp, lengths, 0, 1)
    assert sharded.tobytes() == direct.tobytes()
    shares = pdist.deal(gp, 4, chunk=2)
    full = np.zeros_like(direct)
    for idx in shares:
        full[idx] = engine.anim(gp[idx])
    assert full.tobytes() == direct.tobytes()


def test_generated_clade_set_sample_vs_oracle(engine):
    """The bench's genome generator (synth.make_set_cached, 1 Mb genomes here): a sample of within- and between-clade pairs
    on the engine equals the oracle bit for bit."""
    import tempfile

    from pyani_b200 import synth

    with tempfile.TemporaryDirectory() as td:
        gs = synth.make_set_cached(3, 2, 1_000_000, seed=20260927, workers=3, cache_dir=td)
        base = len(engine)
        gids = [engine.add_genome(np.asarray(g.residues), g.contig_off, g.name) for g in gs]
        og = [oracle.OGenome(np.asarray(g.residues), g.contig_off) for g in gs]
        assert gids[0] == base
        _check(engine, gids, og, [(0, 1), (1, 0), (0, 2), (2, 0), (3, 4), (5, 2)], nofilter=False)


def test_batched_rows_equal_per_pair_rows(engine, small_set):
    """b200ani_anim_rows (all pairs in one pass) == b200ani_anim_alignments pair by pair, in caller order, with and without the filter;
    its result structs == b200ani_anim's."""
    from pyani_b200 import synth

    gs, gids, og = small_set
    pairs = synth.all_directed_pairs(len(gs))[::-1]  # not the internal order
    gp = np.array([[gids[r], gids[q]] for r, q in pairs], dtype=np.int32)
    for nofilter in (False, True):
        res, rows, start = engine.anim_rows(gp, nofilter=nofilter)
        assert res.tobytes() == engine.anim(gp, nofilter=nofilter).tobytes()
        assert start[0] == 0 and start[-1] == len(rows) and (np.diff(start) == res["n_alignments"]).all()
        for i in (0, 1, 7, len(gp) - 1):
            want = engine.anim_alignments(int(gp[i, 0]), int(gp[i, 1]), nofilter=nofilter)
            got = rows[start[i]:start[i + 1]]
            assert got.shape == want.shape and (got == want).all(), (i, nofilter)


def test_exported_and_adopted_index_gives_identical_results(engine, small_set):
    """Seed indexes shared between GPUs (dist.share_indexes), here on one device: genome 0 is staged a second time, the index
    built for the first copy is exported into a torch buffer laid out by dist.index_layout and adopted by the second copy; every
    comparison that uses the second copy as reference or query returns what the first copy returns."""
    import torch

    from pyani_b200 import dist as pdist

    gs, gids, og = small_set
    twin = engine.add_genome(gs[0].residues, gs[0].contig_off, "twin")
    engine.index_build([gids[0]])
    n_mini, bits = engine.index_info(gids[0])
    block_of, off, lbits, block_bytes = pdist.index_layout(np.array([n_mini]), 1)
    assert int(lbits[0]) == bits and block_bytes >= 8 * n_mini + (8 << bits)
    slab = torch.zeros(block_bytes, dtype=torch.uint8, device="cuda:0")
    base = slab.data_ptr()
    engine.index_export(gids[0], *(base + int(o) for o in off[0]))
    engine.sync()
    engine.index_adopt(twin, *(base + int(o) for o in off[0]), n_mini, bits)
    a = engine.anim(np.array([[gids[0], gids[1]], [gids[1], gids[0]], [gids[0], gids[4]]], np.int32))
    b = engine.anim(np.array([[twin, gids[1]], [gids[1], twin], [twin, gids[4]]], np.int32))
    assert a.tobytes() == b.tobytes()
    engine.reset_indexes()   # adopted arrays are forgotten; the next call rebuilds the twin's own index
    del slab
    c = engine.anim(np.array([[twin, gids[1]], [gids[1], twin], [twin, gids[4]]], np.int32))
    assert c.tobytes() == a.tobytes()
