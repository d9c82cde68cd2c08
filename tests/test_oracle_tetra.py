"""The TETRA oracle (oracle/oracle_tetra.c) against the reference's own golden vectors."""
import numpy as np
import pytest

import oracle
from conftest import fasta_text_to_arrays, fixture_genomes


def _z_dict(arr, off):
    c2, c3, c4 = oracle.tetra_counts(arr, off)
    z, obs = oracle.tetra_zscores(c2, c3, c4)
    return {oracle.TETRAMERS[t]: z[t] for t in range(256) if obs[t]}, z, obs


def test_edge_cases_bit_exact(tetra_gold):
    # inline FASTAs (lower case, Ns/IUPAC, multi-record, records shorter than 5) run through the reference
    for name, case in tetra_gold["edge"].items():
        arr, off = fasta_text_to_arrays(case["fasta"])
        got, _, _ = _z_dict(arr, off)
        assert got.keys() == case["z"].keys(), name
        for k, v in case["z"].items():
            assert got[k] == v, (name, k)


@pytest.mark.skipif(not fixture_genomes(), reason="fixture genomes not copied (run tests/golden/make_golden.py)")
def test_fixture_genomes_bit_exact(tetra_gold):
    zs = {}
    for f in fixture_genomes():
        arr, off, _ = oracle.read_fasta(f)
        got, z, _ = _z_dict(arr, off)
        ref = tetra_gold["genomes"][f.stem]
        assert all(got[k] == ref[k] for k in ref) and len(got) == len(ref)
        zs[f.stem] = z
    # tests/test_tetra.py:79-84 of the reference: zscore.json, exact equality
    zj = tetra_gold["zscore_json_NC_002696"]
    assert all(zs["NC_002696"][oracle.TETRAMERS.index(k)] == v for k, v in zj.items())
    # checked-in correlation table (naive left-to-right sums) bit for bit
    orgs = tetra_gold["corr_checked_in"]["orgs"]
    corr = oracle.tetra_corr(np.stack([zs[o] for o in orgs]))
    assert (corr == np.asarray(tetra_gold["corr_checked_in"]["matrix"])).all()
    # live reference under Python >= 3.12 (compensated sum): north_star tolerance 1e-6
    live = np.asarray(tetra_gold["corr_py312_sum"]["matrix"])
    assert tetra_gold["corr_py312_sum"]["orgs"] == orgs
    assert np.abs(corr - live).max() < 1e-12
