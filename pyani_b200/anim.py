"""ANIm on the B200 engine behind the reference's interfaces for the path (pyani/anim.py, pyani_tools.py).

Three levels of drop-in, all thin host code over ``Engine.anim``:

* :func:`parse_result` -- the 4-tuple ``anim.parse_delta`` returns (anim.py:411) from one result struct;
* :func:`calculate_anim` -- in-memory equivalent of ``generate_nucmer_jobs`` + scheduler +
  ``process_deltadir`` (anim.py:155-237,415-497) filling an :class:`ANIResults` exactly like the legacy
  driver does (including its last-file-wins symmetric cells);
* :func:`run_dependency_graph` -- same call shape as ``run_multiprocessing.run_dependency_graph``
  (run_multiprocessing.py:56-84): takes the ``Job`` graph pyani built, recognises the nucmer command
  lines (anim.py:280), runs the comparisons on the GPU and writes minimal ``.delta``/``.filter`` files
  where nucmer / delta-filter would have, so ``subcmd_anim.update_comparison_results``
  (subcmd_anim.py:418-459) and ``process_deltadir`` run unmodified.  Returns 0 like a clean pool run.
"""
from __future__ import annotations

import re
from pathlib import Path
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

from .engine import Engine
from .results import ANIResults  # noqa: F401  (re-exported: the sink type of this module's functions)
from .fasta import read_fasta


class PyaniB200Exception(Exception):
    """Raised where the reference raises PyaniException / PyaniANImException."""


# ------------------------------------------------------------------------------------------------
def parse_result(res) -> Tuple[int, int, float, int]:
    """``(ref_aln_len, qry_aln_len, identity, sim_errs)`` exactly as ``anim.parse_delta`` (anim.py:411).

    An alignment-free comparison raises ``ZeroDivisionError`` like the reference does on an empty
    ``.filter`` (``sum(weighted)/sum(aligned)``, anim.py:396)."""
    if int(res["sum_aligned"]) == 0:
        raise ZeroDivisionError("no alignments between the two genomes (anim.py:396)")
    return (int(res["ref_aln_len"]), int(res["qry_aln_len"]), int(res["sum_weighted"]) / int(res["sum_aligned"]), int(res["sim_errs"]))


def fill_aniresults(org_lengths: Dict[str, int], tuples: Dict[Tuple[str, str], Tuple[int, int, float, int]]) -> ANIResults:
    """``process_deltadir`` (anim.py:415-497) given the per-file 4-tuples instead of files.

    ``tuples[(qname, sname)]`` is the result of the run ``<qname>_vs_<sname>`` (qname = nucmer reference).
    Files are visited in the order of ``sorted(delta_dir.glob("*/*.filter"))`` (anim.py:438), i.e. sorted by
    ``<qname>/<qname>_vs_<sname>.filter``, and both directions write the symmetric cells, so the later
    file wins."""
    results = ANIResults(list(org_lengths.keys()), "ANIm")
    for org, length in org_lengths.items():
        results.add_tot_length(org, org, length)
    # Path objects compare component-wise (directory name first, then file name), which differs from comparing the
    # joined string whenever one stem is a prefix of another followed by a character below '/'
    for qname, sname in sorted(tuples, key=lambda k: (k[0], f"{k[0]}_vs_{k[1]}.filter")):
        if qname not in org_lengths or sname not in org_lengths:
            continue
        q_len, s_len, ident, errs = tuples[(qname, sname)]
        if s_len == 0:
            raise SystemExit("Zero length alignment!")  # anim.py:481-486
        results.add_tot_length(qname, sname, q_len, s_len)
        results.add_sim_errors(qname, sname, errs)
        results.add_pid(qname, sname, ident)
        results.add_coverage(qname, sname, float(q_len) / org_lengths[qname], float(s_len) / org_lengths[sname])
    return results


def calculate_anim(infiles: Sequence[Path], org_lengths: Optional[Dict[str, int]] = None, engine: Optional[Engine] = None,
                   maxmatch: bool = False, nofilter: bool = False) -> ANIResults:
    """All directed comparisons between ``infiles`` on the GPU -> ANIResults (legacy driver semantics,
    average_nucleotide_identity.py:489-575)."""
    eng = engine or Engine(0)
    try:
        files = sorted(Path(f) for f in infiles)  # anim.py:217: sorted file list, both directions
        gids = {f.stem: eng.add_fasta(f) for f in files}
        stems = [f.stem for f in files]
        if org_lengths is None:
            org_lengths = {s: eng.genome_length(gids[s]) for s in stems}
        pairs = [(a, b) for a in stems for b in stems if a != b]
        res = eng.anim(np.array([[gids[a], gids[b]] for a, b in pairs], dtype=np.int32), maxmatch=maxmatch, nofilter=nofilter)
    finally:
        if engine is None:  # an engine created here is not the caller's to clean up
            eng.close()
    tuples = {}
    for (a, b), r in zip(pairs, res):
        tuples[(a, b)] = parse_result(r)
    return fill_aniresults(org_lengths, tuples)


# ------------------------------------------------------------------------------------------------
_NUCMER_RE = re.compile(r"(?:^|\s)(--mum|--maxmatch)\s+-p\s+(\S+)\s+(\S+)\s+(\S+)\s*$")


def _iter_jobs(jobgraph) -> Iterable:
    seen = set()
    stack = list(jobgraph)
    while stack:
        j = stack.pop()
        if id(j) in seen:
            continue
        seen.add(id(j))
        yield j
        stack.extend(getattr(j, "dependencies", []))


def write_minimal_delta(path: Path, ref_path: str, qry_path: str, rows: np.ndarray, ref_ids: List[str], qry_ids: List[str],
                        ref_lens: Sequence[int], qry_lens: Sequence[int]) -> None:
    """A .delta/.filter file holding what ``parse_delta`` and ``nucmer.DeltaIterator`` read: header, one
    '>' line per record pair, 7-field alignment headers, each followed by the terminating ``0`` of an empty
    indel list (format: anim.py:300-325, nucmer.py:317-351).  rows: (ref_rec, qry_rec, rs, re, qs, qe, errs).
    Pure-Python twin of the native emitter ``b200ani_write_delta`` (same bytes; used where no engine is at hand)."""
    path.parent.mkdir(parents=True, exist_ok=True)
    with open(path, "w") as fh:
        fh.write(f"{ref_path} {qry_path}\nNUCMER\n")
        order = np.lexsort((rows[:, 2], rows[:, 1], rows[:, 0])) if len(rows) else []
        cur = None
        for i in order:
            rc, qc, rs, re_, qs, qe, errs = (int(x) for x in rows[i])
            if (rc, qc) != cur:
                fh.write(f">{ref_ids[rc]} {qry_ids[qc]} {ref_lens[rc]} {qry_lens[qc]}\n")
                cur = (rc, qc)
            fh.write(f"{rs} {re_} {qs} {qe} {errs} {errs} 0\n0\n")


def run_dependency_graph(jobgraph, workers: Optional[int] = None, logger=None, engine: Optional[Engine] = None) -> int:
    """GPU stand-in for ``run_multiprocessing.run_dependency_graph(jobgraph, workers, logger) -> int``.

    All nucmer jobs of the graph become ONE batched device pass per mode (``Engine.anim_rows`` with ``all_rows=True``: the
    .delta rows and which of them survive the 1-to-1 filter), then every job's ``.delta`` and ``.filter`` are written where
    nucmer / delta-filter would have put them (native emitter).  ``workers`` is accepted and ignored (one GPU does the work of
    the pool).  Returns 0 on success, the number of comparisons that could not be run otherwise -- the caller's
    ``if cumval > 0: raise PyaniException`` logic (subcmd_anim.py:396-401) keeps working."""
    todo = []
    for job in _iter_jobs(jobgraph):
        cmd = getattr(job, "command", None) or getattr(job, "script", "")
        m = _NUCMER_RE.search(cmd)
        if m and "delta-filter" not in cmd:
            todo.append((m.group(1) == "--maxmatch", m.group(2), m.group(3), m.group(4)))
    if not todo:
        return 0
    eng = engine or Engine(0)
    try:
        files: List[str] = []
        for _, _, f1, f2 in todo:
            for f in (f1, f2):
                if f not in files:
                    files.append(f)
        try:
            if hasattr(eng, "add_fasta_many"):
                gid = dict(zip(files, eng.add_fasta_many(files)))
            else:  # test doubles
                gid = {}
                for f in files:
                    arr, off, ids = read_fasta(f)
                    gid[f] = eng.add_genome(arr, off, Path(f).stem, ids)
        except Exception as exc:
            if logger:
                logger.error("b200ani could not stage the genomes: %s", exc)
            return len(todo)
        failed = 0
        for mode in (False, True):
            jobs = [t for t in todo if t[0] == mode]
            if not jobs:
                continue
            pairs = np.array([[gid[f1], gid[f2]] for _, _, f1, f2 in jobs], dtype=np.int32)
            try:
                _, rows, start, kept = eng.anim_rows(pairs, maxmatch=mode, all_rows=True)
                for i, (_, prefix, f1, f2) in enumerate(jobs):
                    mine, k = rows[start[i]:start[i + 1]], kept[start[i]:start[i + 1]]
                    eng.write_delta(prefix + ".delta", f1, f2, gid[f1], gid[f2], mine)
                    eng.write_delta(prefix + ".filter", f1, f2, gid[f1], gid[f2], mine[k])
            except Exception as exc:  # mirrors non-zero subprocess return codes
                if logger:
                    logger.error("b200ani batch of %d comparisons failed: %s", len(jobs), exc)
                failed += len(jobs)
        return failed
    finally:
        if engine is None:
            eng.close()
