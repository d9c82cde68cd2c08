"""CUDA TETRA path (through the C ABI) vs the oracle and the reference's golden vectors."""
import numpy as np
import pytest

import oracle
from conftest import fasta_text_to_arrays, fixture_genomes

pytestmark = pytest.mark.gpu


def _oracle_all(arr, off):
    c2, c3, c4 = oracle.tetra_counts(arr, off)
    z, obs = oracle.tetra_zscores(c2, c3, c4)
    return c2, c3, c4, z, obs


def test_edge_cases_counts_and_z_bit_exact(engine, tetra_gold):
    for name, case in tetra_gold["edge"].items():
        arr, off = fasta_text_to_arrays(case["fasta"])
        gid = engine.add_genome(arr, off, name)
        c2, c3, c4, z, obs = _oracle_all(arr, off)
        g2, g3, g4 = engine.tetra_counts(gid)
        assert (g2 == c2).all() and (g3 == c3).all() and (g4 == c4).all(), name
        gz, gobs = engine.tetra_zscores([gid])
        assert (gobs[0] == obs).all() and (gz[0] == z).all(), name
        for k, v in case["z"].items():  # and against the reference's own numbers
            assert gz[0][oracle.TETRAMERS.index(k)] == v


def test_random_genomes_match_oracle(engine):
    rng = np.random.default_rng(7)
    gids, zs = [], []
    for i in range(6):
        n_rec = int(rng.integers(1, 5))
        lens = rng.integers(0, 300000, size=n_rec)
        lens[rng.integers(0, n_rec)] = int(rng.integers(200000, 1500000))
        arr = rng.choice(np.frombuffer(b"ACGTacgtNnRY", dtype=np.uint8), size=int(lens.sum()),
                         p=[0.2, 0.27, 0.27, 0.2, 0.01, 0.01, 0.01, 0.01, 0.005, 0.005, 0.005, 0.005])
        off = np.zeros(n_rec + 1, dtype=np.uint64)
        off[1:] = np.cumsum(lens)
        gid = engine.add_genome(arr, off, f"rand{i}")
        c2, c3, c4, z, obs = _oracle_all(arr, off)
        g2, g3, g4 = engine.tetra_counts(gid)
        assert (g2 == c2).all() and (g3 == c3).all() and (g4 == c4).all()
        gids.append(gid)
        zs.append(z)
    gz, gobs = engine.tetra_zscores(gids)
    assert (gz == np.stack(zs)).all()
    corr = engine.tetra_corr(gz)
    assert (corr == oracle.tetra_corr(np.stack(zs))).all()  # serial-order sums: bit-exact
    assert np.allclose(corr, np.corrcoef(gz), atol=1e-12)   # and a real Pearson correlation


@pytest.mark.skipif(not fixture_genomes(), reason="fixture genomes not shipped")
def test_reference_fixture_genomes(engine, tetra_gold):
    """BASELINE.json config[0]: TETRA on tests/test_input -- correlation matrix bit-match."""
    from pyani_b200 import tetra

    zs = tetra.calculate_tetra_zscores(fixture_genomes(), engine=engine)
    for org, ref in tetra_gold["genomes"].items():
        assert zs[org] == ref  # dict equality: same keys, bit-identical floats
    assert all(zs["NC_002696"][k] == v for k, v in tetra_gold["zscore_json_NC_002696"].items())
    corr = tetra.calculate_correlations(zs, engine=engine)
    gold = tetra_gold["corr_checked_in"]
    assert list(corr.index) == gold["orgs"]
    assert (corr.values == np.asarray(gold["matrix"])).all()
    assert np.abs(corr.values - np.asarray(tetra_gold["corr_py312_sum"]["matrix"])).max() < 1e-6


def test_mismatched_tetramer_sets_raise(engine):
    from pyani_b200 import tetra

    with pytest.raises(AssertionError):  # tetra.py:174-175
        tetra.calculate_correlations({"a": {"AAAA": 1.0, "AAAC": 2.0}, "b": {"AAAA": 1.0}}, engine=engine)


def test_fasta_ingest_through_the_pinned_double_buffer(engine, tmp_path):
    """f3: FASTA files -> threaded parse -> pinned double buffer -> pack on the device: a 2.5 Mb genome in pageable memory goes
    through the staging path (counted), its TETRA counts equal the oracle's, and staging the same residues from pinned memory
    (direct copy, not counted) gives the same counts."""
    import torch

    import oracle

    rng = np.random.default_rng(21)
    seq = np.frombuffer(b"ACGTN", np.uint8)[rng.choice(5, size=2_500_000, p=[0.25, 0.25, 0.25, 0.2499, 0.0001])]
    paths = []
    for i, cut in enumerate((1_000_000, 1_700_000)):
        p = tmp_path / f"g{i}.fna"
        with open(p, "wb") as fh:
            fh.write(b">a\n" + bytes(seq[:cut]) + b"\n>b\n" + bytes(seq[cut:]) + b"\n")
        paths.append(p)
    before = engine.staged_bytes()
    gids = engine.add_fasta_many(paths, threads=2)
    assert engine.staged_bytes() - before == 2 * len(seq)
    assert [d["length"] for d in engine.ingested] == [len(seq)] * 2
    off = np.array([0, 1_000_000, len(seq)], np.uint64)
    want = oracle.tetra_counts(seq, off)
    got = engine.tetra_counts(gids[0])
    assert all((a == b).all() for a, b in zip(got, want))
    pinned = torch.empty(len(seq), dtype=torch.uint8, pin_memory=True)
    pinned.numpy()[:] = seq
    mid = engine.staged_bytes()
    g2 = engine.add_genome(pinned.numpy(), off, "pinned")
    assert engine.staged_bytes() == mid
    assert all((a == b).all() for a, b in zip(engine.tetra_counts(g2), want))
