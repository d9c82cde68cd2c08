"""CUDA ANIb path (through the C ABI) vs the CPU oracle: bit-exact on every integer the engine returns."""
import json

import numpy as np
import pytest

import oracle
from conftest import GOLDEN, fixture_genomes

pytestmark = pytest.mark.gpu
FIELDS = ["aln_length", "sim_errs", "pid_milli_sum", "n_frags", "n_pass", "n_hsps", "status"]


@pytest.fixture(scope="module")
def small_set(engine):
    from pyani_b200 import synth

    gs = synth.make_set(3, 2, 300_000, seed=20260926, n_rate=3e-4)
    gids = [engine.add_genome(g.residues, g.contig_off, g.name) for g in gs]
    og = [oracle.OGenome(g.residues, g.contig_off) for g in gs]
    return gs, gids, og


def test_all_pairs_bit_exact(engine, small_set):
    from pyani_b200 import synth

    gs, gids, og = small_set
    pairs = [tuple(p) for p in synth.all_directed_pairs(len(gs))]
    res = engine.anib(np.array([[gids[q], gids[s]] for q, s in pairs], dtype=np.int32))
    for (q, s), got in zip(pairs, res):
        want = oracle.anib_pair(og[q], og[s])[0]
        for f in FIELDS:
            assert int(got[f]) == want[f], (q, s, f, int(got[f]), want[f])
    assert (res["n_pass"] > 0).sum() >= 6
    # other fragment size, single pair
    got = engine.anib(np.array([[gids[0], gids[1]]], dtype=np.int32), fragsize=500)[0]
    want = oracle.anib_pair(og[0], og[1], fragsize=500)[0]
    assert all(int(got[f]) == want[f] for f in FIELDS)


def test_fragment_rows_match(engine, small_set):
    gs, gids, og = small_set
    for q, s in [(0, 1), (1, 0), (2, 0)]:
        want = oracle.anib_pair(og[q], og[s])[1]
        got = engine.anib_fragments(gids[q], gids[s])
        assert got.shape == want.shape and (got == want).all(), (q, s)
        # BLAST row invariants of the reference's fixtures (SURVEY 8c'): length == nident + mismatch + gaps
        assert (got[:, 1] == got[:, 2] + got[:, 3] + got[:, 4]).all() and (got[:, 5] <= 1020).all()


@pytest.mark.skipif(not fixture_genomes(), reason="fixture genomes not shipped")
def test_reference_fixture_genomes(engine):
    """The reference's ANIb fixture genomes: GPU == oracle bit-exact; identity within the reference's own ANIb
    concordance tolerance of its checked-in BLAST+ results (see test_anib_oracle_fixtures)."""
    from pyani_b200 import anib as banib

    gold = json.load(open(GOLDEN / "anib_parse_blast_tab.json"))
    files = fixture_genomes()
    gids, og = [], []
    for f in files:
        arr, off, _ = oracle.read_fasta(f)
        gids.append(engine.add_genome(arr, off, f.stem))
        og.append(oracle.OGenome(arr, off))
    pairs = [(i, j) for i in range(4) for j in range(4) if i != j]
    res = engine.anib(np.array([[gids[i], gids[j]] for i, j in pairs], dtype=np.int32))
    wants = oracle.anib_results_mt(og, pairs)  # host threads: the word stage makes a 4 Mb pair take ~10 s on one core
    for (i, j), r, want in zip(pairs, res, wants):
        for f in FIELDS:
            assert int(r[f]) == want[f], (i, j, f)
        aln, errs, pid = banib.parse_result(r)
        g = gold[f"{files[i].stem}_vs_{files[j].stem}"]
        assert abs(pid - g[2]) < (0.2 if g[2] >= 90 else 5.0)


def test_batched_fragment_rows_equal_per_pair_rows(engine):
    from pyani_b200 import synth

    gs = synth.make_set(2, 2, 200_000, seed=11)
    gids = [engine.add_genome(g.residues, g.contig_off, g.name) for g in gs]
    gp = np.array([[gids[a], gids[b]] for a in range(4) for b in range(4) if a != b][::-1], dtype=np.int32)
    res, rows, start = engine.anib_rows(gp)
    assert res.tobytes() == engine.anib(gp).tobytes()
    assert (np.diff(start) == res["n_pass"]).all()
    for i in (0, 3, len(gp) - 1):
        want = engine.anib_fragments(int(gp[i, 0]), int(gp[i, 1]))
        got = rows[start[i]:start[i + 1]]
        assert got.shape == want.shape and (got == want).all(), i
