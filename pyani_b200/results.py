"""The result sink the legacy driver expects (``pyani_tools.ANIResults``, pyani_tools.py:85-172).

When pyani itself is importable the reference's own class is used, so nothing can drift.  Otherwise (this package
installed without pyani) an array-backed stand-in offers the same attributes and methods: five labelled tables,
``add_tot_length`` / ``add_sim_errors`` / ``add_pid`` / ``add_coverage`` with the reference's symmetric-write rules,
``hadamard``, and the ``data`` property the legacy driver iterates to write its output tables
(average_nucleotide_identity.py:789).
"""
from __future__ import annotations

from typing import Dict, Iterator, List, Optional, Tuple

import numpy as np

_TABLES = ("alignment_lengths", "percentage_identity", "alignment_coverage", "similarity_errors")
# initial cell value of each table (pyani_tools.py:96-105): NaN lengths, zero errors, unit identity / coverage
_INIT = {"alignment_lengths": np.nan, "percentage_identity": 1.0, "alignment_coverage": 1.0, "similarity_errors": 0.0}
# pyani_config.py:60-81: <mode>_<table> in the order ``data`` yields them
_STEM_ORDER = ("alignment_lengths", "percentage_identity", "alignment_coverage", "similarity_errors", "hadamard")


class _ArrayResults:
    def __init__(self, labels: List[str], mode: str) -> None:
        self._labels = list(labels)
        self._pos: Dict[str, int] = {l: i for i, l in enumerate(self._labels)}
        n = len(self._labels)
        self._arr = {t: np.full((n, n), _INIT[t], dtype=np.float64) for t in _TABLES}
        self.zero_error = False
        self.mode = mode

    def _frame(self, a: np.ndarray):
        import pandas as pd

        return pd.DataFrame(a, index=self._labels, columns=self._labels, dtype=float)

    def __getattr__(self, name):  # tables are handed out as DataFrames that write through to the arrays
        if name in _TABLES:
            return self._frame(self.__dict__["_arr"][name])
        raise AttributeError(name)

    def _put(self, table: str, q: str, s: str, v) -> None:
        self._arr[table][self._pos[q], self._pos[s]] = v

    def add_tot_length(self, qname, sname, query_alignment_length, subject_alignment_length=None) -> None:
        self._put("alignment_lengths", qname, sname, query_alignment_length)
        if subject_alignment_length:
            self._put("alignment_lengths", sname, qname, subject_alignment_length)

    def add_sim_errors(self, qname, sname, value, sym: bool = True) -> None:
        self._put("similarity_errors", qname, sname, value)
        if sym:
            self._put("similarity_errors", sname, qname, value)

    def add_pid(self, qname, sname, perc_id) -> None:
        self._put("percentage_identity", qname, sname, perc_id)

    def add_coverage(self, qname, sname, qcover, scover=None) -> None:
        self._put("alignment_coverage", qname, sname, qcover)
        if scover:
            self._put("alignment_coverage", sname, qname, scover)

    @property
    def hadamard(self):
        return self._frame(self._arr["percentage_identity"] * self._arr["alignment_coverage"])

    @property
    def data(self) -> Iterator[Tuple[object, str]]:
        return ((getattr(self, t), f"{self.mode}_{t}") for t in _STEM_ORDER)


def _resolve():
    try:
        from pyani.pyani_tools import ANIResults as ref_cls  # the reference's own sink when pyani is installed

        return ref_cls
    except Exception:
        return _ArrayResults


ANIResults = _resolve()
