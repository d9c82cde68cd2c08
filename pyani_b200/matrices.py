"""Vectorised assembly of the five N x N result matrices of a v0.3 run.

Same cells as ``pyani_orm.update_comparison_matrices`` (pyani_orm.py:618-666) fills one ``.loc`` at a time:
rows = pyani query (= nucmer reference) genome, columns = subject; float64 throughout; diagonals
identity 1, coverage 1, aln_lengths = genome length, sim_errors 0, hadamard 1; cells without a comparison NaN.
``Comparison`` fields as ``subcmd_anim.update_comparison_results`` derives them (subcmd_anim.py:438-455):
aln_length = parse_delta()[0], cov_query = that / len(query).
"""
from __future__ import annotations

from typing import Dict, Sequence

import numpy as np


def anim_matrices(genome_ids: Sequence, lengths: Sequence[int], pairs: np.ndarray, results: np.ndarray) -> Dict[str, "object"]:
    """pairs[i] = (index of the query/nucmer-reference genome, index of the subject) into ``genome_ids``;
    results = structured array from ``Engine.anim``.  Pairs without alignments stay NaN (the reference would have
    aborted on them, anim.py:396)."""
    import pandas as pd

    order = np.argsort(np.asarray(genome_ids))  # sorted genome ids label rows and columns (pyani_orm.py:628)
    rank = np.empty(len(order), dtype=np.int64)
    rank[order] = np.arange(len(order))
    ids = [genome_ids[i] for i in order]
    n = len(ids)
    L = np.asarray(lengths, dtype=np.float64)
    mats = {k: np.full((n, n), np.nan) for k in ("identity", "coverage", "aln_lengths", "sim_errors", "hadamard")}
    for k, v in (("identity", 1.0), ("coverage", 1.0), ("sim_errors", 0.0), ("hadamard", 1.0)):
        np.fill_diagonal(mats[k], v)
    mats["aln_lengths"][rank, rank] = L
    p = np.asarray(pairs, dtype=np.int64).reshape(-1, 2)
    ok = results["sum_aligned"] > 0
    q, s = rank[p[ok, 0]], rank[p[ok, 1]]
    ident = results["sum_weighted"][ok] / results["sum_aligned"][ok]
    cov = results["ref_aln_len"][ok] / L[p[ok, 0]]
    mats["identity"][q, s] = ident
    mats["coverage"][q, s] = cov
    mats["aln_lengths"][q, s] = results["ref_aln_len"][ok]
    mats["sim_errors"][q, s] = results["sim_errs"][ok]
    mats["hadamard"][q, s] = ident * cov
    return {k: pd.DataFrame(v, index=ids, columns=ids, dtype=float) for k, v in mats.items()}


def comparison_rows(genome_ids: Sequence, lengths: Sequence[int], pairs: np.ndarray, results: np.ndarray, version: str,
                    maxmatch: bool = False) -> list:
    """One dict per directed comparison with the columns of ``pyani_orm.Comparison`` (pyani_orm.py:262-322), valued as
    ``subcmd_anim.update_comparison_results`` values them (subcmd_anim.py:438-455): ``aln_length`` = aligned length on the
    query (= nucmer reference), ``cov_query`` / ``cov_subject`` = aligned length over genome length, ``identity`` =
    sum_weighted / sum_aligned, ``program = "b200ani"`` (a cache key of its own in ``filter_existing_comparisons``,
    pyani_orm.py:416-491).  Ready for ``session.bulk_insert_mappings(Comparison, rows)``; pairs without
    any alignment are left out (the reference aborts on them, anim.py:396,481-486)."""
    p = np.asarray(pairs, dtype=np.int64).reshape(-1, 2)
    L = np.asarray(lengths, dtype=np.float64)
    rows = []
    for (q, s), r in zip(p, results):
        if int(r["sum_aligned"]) <= 0:
            continue
        rows.append({"query_id": genome_ids[q], "subject_id": genome_ids[s], "aln_length": int(r["ref_aln_len"]),
                     "sim_errs": int(r["sim_errs"]), "identity": int(r["sum_weighted"]) / int(r["sum_aligned"]),
                     "cov_query": int(r["ref_aln_len"]) / L[q], "cov_subject": int(r["qry_aln_len"]) / L[s],
                     "program": "b200ani", "version": version, "fragsize": None, "maxmatch": bool(maxmatch),
                     "kmersize": None, "minmatch": None})
    return rows
