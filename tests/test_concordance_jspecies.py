"""The reference's own concordance targets: JSpecies output for three Corynebacterium genomes
(tests/fixtures/concordance/jspecies_output.tab, compared in tests/test_concordance.py:169-302 with
tolerances TETRA 0.1, ANIm 0.1, ANIb 0.2 above 90 % identity and 5 below).

The JSpecies table is committed as tests/golden/jspecies_concordance.json; the three genomes (9 MB) are copied to
oracle/_ref/fixtures/concordance by tests/golden/make_golden.py (git-ignored, shipped to the GPU box).

Where this engine stands against those targets (CPU oracle == CUDA engine, bit for bit):
  TETRA  within 1e-3 of JSpecies                                   (reference tolerance 0.1: met)
  ANIm   98.31 vs 98.19, 84.23 vs 84.11/84.09, 84.35 vs 84.53/84.55  (reference tolerance 0.1: NOT met, worst 0.20)
  ANIb   98.25..98.30 vs 98.17/98.19                                (tolerance 0.2: met)
         73.5 vs 68.7..68.8 on the two distant pairs                (tolerance 5: met, +4.7..+4.85.  The reference keeps that
         tolerance wide because JSpecies ran legacy BLAST and "the results can be quite different" from BLAST+ below 90 %,
         test_concordance.py:119-127; with the 11-mer word anchors 25 % more fragments pass here than with unique exact
         seeds >= 14 alone, at the same mean identity)
The assertions below pin the current deviations (so they can only shrink) and state the reference tolerance next
to each; the unmet one (ANIm) is also written as an xfail test so the gap shows up in every test report.
"""
import json

import numpy as np
import pytest

import oracle
from conftest import GOLDEN, fixture_genomes

FILES = fixture_genomes("concordance")
pytestmark = pytest.mark.skipif(len(FILES) != 3, reason="concordance genomes not copied (tests/golden/make_golden.py)")
JS = json.load(open(GOLDEN / "jspecies_concordance.json"))


def target(method, a, b):
    """Both JSpecies cells of the unordered pair (a, b): the table is not symmetric for ANIb/ANIm."""
    t = JS[method]
    i, j = t["rows"].index(a), t["names"].index(b)
    i2, j2 = t["rows"].index(b), t["names"].index(a)
    return t["values"][i][j], t["values"][i2][j2]


def worst_dev(value, method, a, b):
    return max(abs(value - x) for x in target(method, a, b))


@pytest.fixture(scope="module")
def arrays():
    out = {}
    for f in FILES:
        a, off, _ = oracle.read_fasta(f)
        out[f.name] = (a, off)
    return out


@pytest.fixture(scope="module")
def ogenomes(arrays):
    return {n: oracle.OGenome(a, off) for n, (a, off) in arrays.items()}


def anim_devs(identity_of):
    devs = {}
    names = [f.name for f in FILES]
    for r in names:
        for q in names:
            if r != q:
                devs[(r, q)] = worst_dev(100 * identity_of(r, q), "ANIm", r, q)
    return devs


def anib_devs(pid_of):
    hi, lo = {}, {}
    names = [f.name for f in FILES]
    for q in names:
        for s in names:
            if q != s:
                t = target("ANIb", q, s)
                (hi if min(t) > 90 else lo)[(q, s)] = worst_dev(pid_of(q, s), "ANIb", q, s)
    return hi, lo


# ------------------------------------------------------------------------------------------ CPU oracle
def test_tetra_oracle(arrays):
    names = [f.name for f in FILES]
    z = np.array([oracle.tetra_zscores(*oracle.tetra_counts(*arrays[n]))[0] for n in names])
    corr = oracle.tetra_corr(z)
    for i, a in enumerate(names):
        for j, b in enumerate(names):
            if i != j:
                assert worst_dev(corr[i, j], "Tetra", a, b) < 1e-3  # reference tolerance: 0.1 (test_concordance.py:163-165)


def oracle_identity(ogenomes):
    cache = {}

    def f(r, q):
        if (r, q) not in cache:
            res, _, _ = oracle.anim_pair(ogenomes[r], ogenomes[q])
            cache[(r, q)] = res["sum_weighted"] / res["sum_aligned"]
        return cache[(r, q)]

    return f


def test_anim_oracle_pinned(ogenomes):
    devs = anim_devs(oracle_identity(ogenomes))
    # reference tolerance: 0.1 (test_concordance.py:157-159): met on the 98 % pair (0.05) and one 84 % pair (0.08 / 0.10),
    # not on the other (-0.20 / -0.22; 45 alignments, 37 kb aligned -- JSpecies itself puts two genomes that are 98.2 %
    # identical to each other 0.44 points apart against the third)
    assert max(devs.values()) < 0.23, devs
    assert sorted(devs.values())[3] < 0.102, devs


@pytest.mark.xfail(reason="ANIm is within 0.1 of JSpecies on 4 of the 6 directed pairs of this set, 0.22 on the other two", strict=False)
def test_anim_oracle_reference_tolerance(ogenomes):
    devs = anim_devs(oracle_identity(ogenomes))
    assert max(devs.values()) < 0.1, devs


def oracle_pid(ogenomes):
    def f(q, s):
        res, _, _ = oracle.anib_pair(ogenomes[q], ogenomes[s])
        return res["pid_milli_sum"] / max(res["n_pass"], 1) / 1000

    return f


def test_anib_oracle_pinned(ogenomes):
    hi, lo = anib_devs(oracle_pid(ogenomes))
    assert len(hi) == 2 and len(lo) == 4
    assert max(hi.values()) < 0.2, hi  # reference tolerance above 90 % identity: 0.2 (test_concordance.py:131-141) -- met
    assert max(lo.values()) < 5.0, lo  # reference tolerance below: 5 (test_concordance.py:143-154) -- met (4.7)


# ------------------------------------------------------------------------------------------ CUDA engine
@pytest.fixture(scope="module")
def gids(engine, arrays):
    return {n: engine.add_genome(a, off, n) for n, (a, off) in arrays.items()}


@pytest.mark.gpu
def test_concordance_gpu(engine, arrays, ogenomes, gids):
    names = [f.name for f in FILES]
    pairs = [(r, q) for r in names for q in names if r != q]
    gp = np.array([[gids[r], gids[q]] for r, q in pairs], np.int32)
    # TETRA
    z, _ = engine.tetra_zscores([gids[n] for n in names])
    corr = engine.tetra_corr(z)
    for i, a in enumerate(names):
        for j, b in enumerate(names):
            if i != j:
                assert worst_dev(corr[i, j], "Tetra", a, b) < 1e-3
    # ANIm: the engine returns exactly the oracle's integers, so the same deviations from JSpecies
    res = engine.anim(gp)
    ident = {p: int(x["sum_weighted"]) / int(x["sum_aligned"]) for p, x in zip(pairs, res)}
    for (r, q), x in zip(pairs, res):
        want, _, _ = oracle.anim_pair(ogenomes[r], ogenomes[q])
        for f in ("ref_aln_len", "qry_aln_len", "sum_aligned", "sum_weighted", "sim_errs", "n_alignments"):
            assert int(x[f]) == want[f], (r, q, f)
    assert max(anim_devs(lambda r, q: ident[(r, q)]).values()) < 0.23
    # ANIb
    resb = engine.anib(gp)
    pid = {p: int(x["pid_milli_sum"]) / max(int(x["n_pass"]), 1) / 1000 for p, x in zip(pairs, resb)}
    wants = oracle.anib_results_mt([ogenomes[n] for n in names], [(names.index(q), names.index(s)) for q, s in pairs])
    for (q, s), x, want in zip(pairs, resb, wants):
        for f in ("aln_length", "sim_errs", "pid_milli_sum", "n_frags", "n_pass"):
            assert int(x[f]) == want[f], (q, s, f)
    hi, lo = anib_devs(lambda q, s: pid[(q, s)])
    assert max(hi.values()) < 0.2 and max(lo.values()) < 5.0


# ------------------------------------------------------------------------------------------ second JSpecies table
# tests/test_JSpecies/jspecies_results.tab: JSpecies 1.2.1 on the four Caulobacter genomes of the reference's ANIm / ANIb
# fixtures (tests/test_JSpecies/README.md).  Here every tolerance of the reference's concordance tests is met:
# ANIm within 0.068 (0.1), ANIb 0.008 above 90 % (0.2) and 2.9 below (5), TETRA 5e-6 (0.1).
JS4 = json.load(open(GOLDEN / "jspecies_caulobacter.json"))
CAULO = fixture_genomes()


def target4(method, a, b):
    t = JS4[method]
    return t["values"][t["rows"].index(a)][t["names"].index(b)], t["values"][t["rows"].index(b)][t["names"].index(a)]


@pytest.mark.skipif(len(CAULO) != 4, reason="fixture genomes not copied (tests/golden/make_golden.py)")
def test_caulobacter_table_oracle():
    arrays = {f.name: oracle.read_fasta(f)[:2] for f in CAULO}
    og = {n: oracle.OGenome(a, off) for n, (a, off) in arrays.items()}
    names = list(arrays)
    z = np.array([oracle.tetra_zscores(*oracle.tetra_counts(*arrays[n]))[0] for n in names])
    corr = oracle.tetra_corr(z)
    worst = {"anim": 0.0, "anib_hi": 0.0, "anib_lo": 0.0, "tetra": 0.0}
    bpairs = [(i, j) for i in range(len(names)) for j in range(len(names)) if i != j]
    anib = dict(zip(bpairs, oracle.anib_results_mt([og[n] for n in names], bpairs)))
    for i, a in enumerate(names):
        for j, b in enumerate(names):
            if a == b:
                continue
            res, _, _ = oracle.anim_pair(og[a], og[b])
            idy = 100 * res["sum_weighted"] / res["sum_aligned"]
            worst["anim"] = max(worst["anim"], max(abs(idy - x) for x in target4("ANIm", a, b)))
            rb = anib[(i, j)]
            pid = rb["pid_milli_sum"] / rb["n_pass"] / 1000
            tb = target4("ANIb", a, b)
            k = "anib_hi" if min(tb) > 90 else "anib_lo"
            worst[k] = max(worst[k], max(abs(pid - x) for x in tb))
            worst["tetra"] = max(worst["tetra"], max(abs(corr[i, j] - x) for x in target4("Tetra", a, b)))
    assert worst["anim"] < 0.08, worst    # measured 0.068; tests/test_concordance.py:157-159 allows 0.1
    assert worst["anib_hi"] < 0.2, worst  # :131-141
    assert worst["anib_lo"] < 5.0, worst  # :143-154
    assert worst["tetra"] < 1e-4, worst   # :163-165 allows 0.1
