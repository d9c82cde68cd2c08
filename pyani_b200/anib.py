"""ANIb on the B200 engine behind the reference's interfaces for the path (pyani/anib.py).

* :func:`get_fraglength_dict` / :func:`fragment_lengths` -- what ``anib.fragment_fasta_files`` /
  ``get_fraglength_dict`` (anib.py:164-232) produce, without writing fragment FASTA files: the engine fragments
  by index arithmetic on the staged genome (1020-nt pieces per record, trailing short piece kept, ``frag%05d``
  numbered across records from 1).
* :func:`parse_result` -- the 3-tuple of ``anib.parse_blast_tab`` (anib.py:667) from one result struct.
* :func:`calculate_anib` / :func:`fill_aniresults` -- ``anib.process_blast`` (anib.py:496-565): lengths, identity
  (= 0.01 * mean_pid) and coverage (= aln_length / query genome length) fill the ``[query, subject]`` cell only;
  similarity errors are written symmetrically (anib.py:562).
"""
from __future__ import annotations

import re
from pathlib import Path
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from .results import ANIResults
from .engine import Engine

FRAGSIZE = 1020  # pyani_config.FRAGSIZE (pyani_config.py:95)


def fragment_lengths(record_lengths: Sequence[int], fragsize: int = FRAGSIZE) -> Dict[str, int]:
    """{fragment id: length} exactly as anib.fragment_fasta_files numbers and cuts them (anib.py:192-200)."""
    out, idx = {}, 0
    for L in record_lengths:
        pos = 0
        while pos < L:
            idx += 1
            out["frag%05d" % idx] = min(fragsize, L - pos)
            pos += fragsize
    return out


def get_fraglength_dict(fastafiles: Sequence[Path], fragsize: int = FRAGSIZE) -> Dict[str, Dict[str, int]]:
    """anib.get_fraglength_dict (anib.py:207-221) for original (unfragmented) FASTA files."""
    from .fasta import read_fasta

    out = {}
    for f in fastafiles:
        _, off, _ = read_fasta(f)
        out[Path(f).stem] = fragment_lengths(np.diff(off.astype(np.int64)).tolist(), fragsize)
    return out


def parse_result(res) -> Tuple[int, int, float]:
    """``(aln_length, sim_errors, mean_pid)`` as anib.parse_blast_tab (anib.py:661-667); (0, 0, 0) if nothing passes."""
    n = int(res["n_pass"])
    mean_pid = (int(res["pid_milli_sum"]) / n / 1000.0) if n else 0.0
    return int(res["aln_length"]), int(res["sim_errs"]), mean_pid


def fill_aniresults(org_lengths: Dict[str, int], tuples: Dict[Tuple[str, str], Tuple[int, int, float]], mode: str = "ANIb") -> ANIResults:
    """anib.process_blast (anib.py:496-565) given the per-file 3-tuples instead of .blast_tab files."""
    results = ANIResults(list(org_lengths.keys()), mode)
    for org, length in org_lengths.items():
        results.add_tot_length(org, org, length)
    # files are visited in sorted(<q>_vs_<s>.blast_tab) order (pyani_files.get_input_files, pyani_files.py:124);
    # add_sim_errors is called with its default sym=True (anib.py:562), so that table is symmetric and the later
    # file wins, while lengths, identity and coverage only fill the [query, subject] cell
    for (qname, sname) in sorted(tuples, key=lambda k: f"{k[0]}_vs_{k[1]}.blast_tab"):
        if qname not in org_lengths or sname not in org_lengths:
            continue
        aln, errs, pid = tuples[(qname, sname)]
        results.add_tot_length(qname, sname, aln)
        results.add_sim_errors(qname, sname, errs)
        results.add_pid(qname, sname, 0.01 * pid)
        results.add_coverage(qname, sname, float(aln) / org_lengths[qname])
    return results


def calculate_anib(infiles: Sequence[Path], org_lengths: Optional[Dict[str, int]] = None, engine: Optional[Engine] = None,
                   fragsize: int = FRAGSIZE) -> ANIResults:
    """All directed comparisons between ``infiles`` on the GPU -> ANIResults (unified_anib semantics,
    average_nucleotide_identity.py:694-760)."""
    eng = engine or Engine(0)
    try:
        files = sorted(Path(f) for f in infiles)
        gids = {f.stem: eng.add_fasta(f) for f in files}
        stems = [f.stem for f in files]
        if org_lengths is None:
            org_lengths = {s: eng.genome_length(gids[s]) for s in stems}
        pairs = [(a, b) for a in stems for b in stems if a != b]
        res = eng.anib(np.array([[gids[a], gids[b]] for a, b in pairs], dtype=np.int32), fragsize=fragsize)
    finally:
        if engine is None:
            eng.close()
    return fill_aniresults(org_lengths, {p: parse_result(r) for p, r in zip(pairs, res)})


def write_minimal_blast_tab(path: Path, sname: str, rows: np.ndarray, slen: int = 0) -> None:
    """A ``.blast_tab`` file holding what ``anib.parse_blast_tab`` reads (anib.py:609-665): one line per kept HSP with the
    15 columns of the reference's ``-outfmt`` (anib.py:468-469): qseqid sseqid length mismatch pident nident qlen slen
    qstart qend sstart send positive ppos gaps.  rows: ``Engine.anib_fragments`` output (fragment index, length, mismatch,
    gaps, nident, qlen, score, round(1e5 * nident / length)).  Fragment ids follow ``anib.fragment_fasta_files``
    (``frag%05d``, numbered from 1 across the records of the file, anib.py:192-200); the coordinate columns, which the
    reference never reads, are written as 0."""
    path = Path(path)
    path.parent.mkdir(parents=True, exist_ok=True)
    with open(path, "w") as fh:
        for frag, length, mism, gaps, nident, qlen, _score, pid_milli in np.asarray(rows, dtype=np.int64).reshape(-1, 8):
            pid = f"{pid_milli / 1000:.3f}"
            fh.write("\t".join(str(x) for x in (f"frag{frag + 1:05d}", sname, length, mism, pid, nident, qlen, slen, 0, 0, 0, 0, nident, pid, gaps)) + "\n")


_BLASTN_RE = re.compile(r"-out\s+(\S+)\s+-query\s+(\S+)\s+-db\s+(\S+)\s.*-task\s+blastn\s*$")
_MAKEDB_RE = re.compile(r"-dbtype\s+nucl\s+-in\s+(\S+)\s+-title\s+\S+\s+-out\s+(\S+)\s*$")


def run_dependency_graph(jobgraph, workers: Optional[int] = None, logger=None, engine: Optional[Engine] = None, fragsize: int = FRAGSIZE) -> int:
    """GPU stand-in for ``run_multiprocessing.run_dependency_graph`` on an ANIb job graph (``anib.make_job_graph``,
    anib.py:296-347): the ``blastn`` jobs (command lines of ``construct_blastn_cmdline``, anib.py:451-471) are run on the
    engine and their ``.blast_tab`` files written where BLAST+ would have put them; the ``makeblastdb`` dependencies
    (``construct_makeblastdb_cmd``, anib.py:378-391) are only read for the database-name -> input-FASTA mapping.  The query
    genome is fragmented on the device, so the ``-fragments`` FASTA named by ``-query`` is not read either.
    Returns 0 on success, the number of failed comparisons otherwise (a summed return code)."""
    from .anim import _iter_jobs
    from .fasta import read_fasta

    db_to_fasta: Dict[str, str] = {}
    todo = []
    for job in _iter_jobs(jobgraph):
        cmd = getattr(job, "command", None) or getattr(job, "script", "")
        m = _MAKEDB_RE.search(cmd)
        if m:
            db_to_fasta[m.group(2)] = m.group(1)
            continue
        m = _BLASTN_RE.search(cmd)
        if m:
            todo.append((m.group(1), m.group(2), m.group(3)))
    if not todo:
        return 0
    by_name = {Path(db).name: fa for db, fa in db_to_fasta.items()}
    eng = engine or Engine(0)
    try:
        jobs, failed = [], 0
        for out, query, db in todo:
            try:
                jobs.append((out, by_name[Path(query).name.replace("-fragments", "")], db_to_fasta[db]))
            except KeyError as exc:
                if logger:
                    logger.error("b200ani ANIb comparison %s vs %s: unknown database / query %s", query, db, exc)
                failed += 1
        files: List[str] = []
        for _, q, s_ in jobs:
            for f in (q, s_):
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
            pairs = np.array([[gid[q], gid[s_]] for _, q, s_ in jobs], dtype=np.int32).reshape(-1, 2)
            _, rows, start = eng.anib_rows(pairs, fragsize=fragsize)  # ONE batched device pass for every blastn job
            for i, (out, q, s_) in enumerate(jobs):
                write_minimal_blast_tab(Path(out), Path(s_).stem, rows[start[i]:start[i + 1]], slen=eng.genome_length(gid[s_]))
        except Exception as exc:  # mirrors non-zero subprocess return codes
            if logger:
                logger.error("b200ani ANIb batch of %d comparisons failed: %s", len(jobs), exc)
            failed += len(jobs)
        return failed
    finally:
        if engine is None:
            eng.close()
