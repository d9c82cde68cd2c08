"""The ANIb-equivalent aligner oracle against the reference's checked-in BLAST+ results.

blastn cannot run here; the pin is the 12 .blast_tab files of the reference's fixtures reduced by the reference's
own parse_blast_tab (tests/golden/anib_parse_blast_tab.json).  Tolerance = the reference's own concordance
tolerance for ANIb (tests/test_concordance.py:118-153,250-253): 0.2 percentage points where identity >= 90 %,
5 points below.  Measured: |delta| <= 0.005 for the 99.99 % pair, +0.5 .. +0.95 points for the 80-84 % pairs, aligned
length 1.7-4.5 % below BLAST's there (before the 11-mer word anchors: +0.7 .. +1.3 points, 4-10 % short).
"""
import json

import pytest

import oracle
from conftest import GOLDEN, fixture_genomes

pytestmark = pytest.mark.skipif(not fixture_genomes(), reason="fixture genomes not copied (tests/golden/make_golden.py)")


def test_anib_identity_within_reference_tolerance():
    gold = json.load(open(GOLDEN / "anib_parse_blast_tab.json"))
    g = {}
    for f in fixture_genomes():
        a, off, _ = oracle.read_fasta(f)
        g[f.stem] = oracle.OGenome(a, off)
    names = sorted(g)
    pairs = [(names.index(n.split("_vs_")[0]), names.index(n.split("_vs_")[1])) for n in gold]
    results = oracle.anib_results_mt([g[n] for n in names], pairs)
    for (name, (aln, errs, pid)), res in zip(gold.items(), results):
        mean_pid = res["pid_milli_sum"] / res["n_pass"] / 1000
        tol = 0.2 if pid >= 90 else 5.0
        assert abs(mean_pid - pid) < tol, (name, mean_pid, pid)
        assert 0.95 * aln < res["aln_length"] < 1.02 * aln, (name, res["aln_length"], aln)
        assert abs(mean_pid - pid) < 1.0
        if pid >= 99:
            assert abs(res["aln_length"] - aln) < 2500 and res["n_frags"] in (3940, 3964)
    # per-fragment rows of one pair (single-threaded entry point): the threaded driver returns the same sums
    q, s = pairs[0]
    res, rows, _ = oracle.anib_pair(g[names[q]], g[names[s]])
    assert res == results[0]
    assert res["n_pass"] == len(rows) and (rows[:, 1] == rows[:, 2] + rows[:, 3] + rows[:, 4]).all()
