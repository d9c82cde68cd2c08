"""parse_delta restatement (oracle/oracle_delta.c) against tuples computed by the reference itself."""
import json

import numpy as np

import oracle
from conftest import GOLDEN


def test_parse_delta_tuples_exact():
    gold = json.load(open(GOLDEN / "anim_parse_delta.json"))
    arrays = np.load(GOLDEN / "anim_alignments.npz")
    assert len(gold) == 25
    for name, (ral, qal, ident, errs) in gold.items():
        got = oracle.parse_delta_reduce(arrays[name].astype(np.int64))
        assert got == (ral, qal, ident, errs), name
    # the value asserted by the reference's own test (tests/test_anim.py:99)
    assert tuple(gold["test.delta"]) == (4016947, 4017751, 0.9994621994447228, 2191)


def test_touching_and_duplicate_intervals():
    # anim.py:399-409 with merge_overlaps(strict=False): [1,10] and [10,20] merge -> 20; duplicates collapse
    rows = np.array([[0, 0, 1, 10, 1, 10, 0], [0, 0, 10, 20, 30, 40, 1], [0, 0, 10, 20, 40, 30, 1],
                     [0, 1, 22, 30, 5, 13, 2]], dtype=np.int64)
    ral, qal, ident, errs = oracle.parse_delta_reduce(rows)
    assert ral == 20 + 9 and qal == 10 + 11 + 9 and errs == 4
    assert ident == (20 + 22 + 22 + 18 - 8) / (20 + 22 + 22 + 18)
