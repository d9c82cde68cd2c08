"""parse_blast_tab restatement (oracle.parse_blast_tab_reduce) against the reference's own results on its
checked-in BLAST+ output (tests/test_anib.py:361-402 of the reference)."""
import json

import numpy as np
import pytest

import oracle
from conftest import GOLDEN


def test_parse_blast_tab_tuples():
    gold = json.load(open(GOLDEN / "anib_parse_blast_tab.json"))
    rows = np.load(GOLDEN / "anib_blast_rows.npz")
    assert len(gold) == 12
    for name, (aln, errs, pid) in gold.items():
        g_aln, g_errs, g_pid = oracle.parse_blast_tab_reduce(rows[name])
        assert (g_aln, g_errs) == (aln, errs), name
        assert g_pid == pytest.approx(pid, rel=1e-13), name
    # the documented known answer (tests/test_anib.py:387-392): (4016551, 93, 99.997693577050029)
    assert gold["NC_002696_vs_NC_011916"][:2] == [4016551, 93]
    assert gold["NC_002696_vs_NC_011916"][2] == pytest.approx(99.997693577050029, rel=1e-12)


def test_filter_is_strict_and_first_row_wins():
    rows = np.array([
        [1, 714, 0, 100.0, 1020, 0],    # coverage exactly 0.7 -> dropped (strict >, anib.py:646)
        [1, 800, 10, 98.75, 1020, 0],   # first kept row of fragment 1
        [1, 900, 0, 100.0, 1020, 0],    # later row of the same fragment -> ignored
        [2, 1000, 700, 30.0, 1020, 0],  # ani_pid = 300/1020 < 0.3 -> dropped
        [3, 1030, 20, 97.0, 1020, 12],  # gaps are removed from the aligned length
    ], dtype=np.float64)
    assert oracle.parse_blast_tab_reduce(rows) == (800 + 1018, 10 + 32, (98.75 + 97.0) / 2)
    assert oracle.parse_blast_tab_reduce(np.zeros((0, 6))) == (0, 0, 0.0)
