"""Host-side mirror of the reference's ANIm interfaces (pyani_b200/anim.py) -- no GPU needed."""
import io
import json

import numpy as np
import pandas as pd
import pytest

from conftest import GOLDEN
from pyani_b200 import anim as banim
from pyani_b200.engine import ANIM_DTYPE


def _gold():
    tuples = json.load(open(GOLDEN / "anim_parse_delta.json"))
    legacy = json.load(open(GOLDEN / "anim_process_deltadir.json"))
    return tuples, legacy


def test_fill_aniresults_matches_reference_process_deltadir():
    """anim.process_deltadir (run by the reference itself on its fixture directory) vs fill_aniresults fed the
    per-file tuples: every cell of all four tables, including the last-file-wins symmetric cells (A.2)."""
    tuples, legacy = _gold()
    per_file = {}
    for name, t in tuples.items():
        if name.endswith(".filter"):
            q, s = name[: -len(".filter")].split("_vs_")
            per_file[(q, s)] = tuple(t)
    res = banim.fill_aniresults(legacy["org_lengths"], per_file)
    for tab in ("alignment_lengths", "similarity_errors", "percentage_identity", "alignment_coverage"):
        g = legacy[tab]
        want = pd.DataFrame(g["values"], index=g["index"], columns=g["columns"])
        got = getattr(res, tab).loc[g["index"], g["columns"]]
        assert (got.values == want.values).all(), tab
    # the reference's own asserted table (tests/test_anim.py:243-255)
    csv = pd.read_csv(io.StringIO(legacy["deltadir_result_csv"]), index_col=0)
    got = res.percentage_identity.loc[csv.index, csv.columns]
    assert np.allclose(got.values, csv.values, atol=5e-7)
    # Appendix A.2
    assert res.alignment_lengths.loc["NC_002696", "NC_010338"] == 1385260
    assert res.similarity_errors.loc["NC_002696", "NC_010338"] == 206821


def test_parse_result_tuple_and_empty():
    r = np.zeros(1, dtype=ANIM_DTYPE)[0]
    r["ref_aln_len"], r["qry_aln_len"], r["sum_aligned"], r["sum_weighted"], r["sim_errs"] = 10, 11, 33, 23, 5
    assert banim.parse_result(r) == (10, 11, 23 / 33, 5)  # the docstring example of anim.py:329-344
    with pytest.raises(ZeroDivisionError):
        banim.parse_result(np.zeros(1, dtype=ANIM_DTYPE)[0])


class _FakeEngine:
    """Stands in for Engine so the job-graph backend can be tested without a GPU."""

    def __init__(self):
        self.n = 0
        self.calls = []
        self.recs = {}

    def add_genome(self, arr, off, name, ids):
        self.recs[self.n] = (ids, np.diff(np.asarray(off).astype(np.int64)))
        self.n += 1
        return self.n - 1

    def anim_rows(self, pairs, maxmatch=False, nofilter=False, all_rows=False):
        self.calls.append((np.asarray(pairs).tolist(), maxmatch, nofilter, all_rows))
        per_pair = np.array([[0, 0, 1, 10, 1, 11, 5], [0, 0, 15, 20, 30, 25, 0]], dtype=np.int64)
        m = len(pairs)
        rows = np.concatenate([per_pair] * m) if m else np.zeros((0, 7), np.int64)
        start = np.arange(m + 1, dtype=np.int64) * 2
        kept = np.tile([True, False], m)
        return None, rows, start, kept

    def write_delta(self, path, ref_path, qry_path, ref_gid, qry_gid, rows):
        from pathlib import Path

        (ri, rl), (qi, ql) = self.recs[ref_gid], self.recs[qry_gid]
        banim.write_minimal_delta(Path(path), ref_path, qry_path, np.asarray(rows).reshape(-1, 7), ri, qi, rl, ql)


class _Job:  # shape of pyani_jobs.Job (pyani_jobs.py:76-120)
    def __init__(self, name, command):
        self.name, self.command, self.dependencies = name, command, []

    def add_dependency(self, job):
        self.dependencies.append(job)


def test_run_dependency_graph_writes_files_parse_delta_can_read(tmp_path):
    import oracle

    fa, fb = tmp_path / "A.fna", tmp_path / "B.fna"
    fa.write_text(">a1 d\nACGTACGTAC\nGTACGTAC\n")
    fb.write_text(">b1 d\nACGTACGTACGTAC\n>b2\nACGT\n")
    out = tmp_path / "out" / "nucmer_output" / "A" / "A_vs_B"
    # command strings exactly as anim.construct_nucmer_cmdline builds them (anim.py:280-288)
    njob = _Job("t_000000-n", f"nucmer --mum -p {out} {fa} {fb}")
    fjob = _Job("t_000000-f", f"delta_filter_wrapper.py delta-filter -1 {out}.delta {out}.filter")
    fjob.add_dependency(njob)
    eng = _FakeEngine()
    assert banim.run_dependency_graph([fjob], workers=4, engine=eng) == 0
    assert len(eng.calls) == 1 and eng.calls[0][0] == [[0, 1]] and eng.calls[0][3] is True  # ONE batched pass for the whole graph
    d = oracle.parse_delta_text(str(out) + ".delta")
    f = oracle.parse_delta_text(str(out) + ".filter")
    assert d[:, 2:].tolist() == [[1, 10, 1, 11, 5], [15, 20, 30, 25, 0]] and f[:, 2:].tolist() == [[1, 10, 1, 11, 5]]
    text = open(str(out) + ".filter").read().splitlines()
    assert text[1] == "NUCMER" and text[2] == ">a1 b1 18 14" and text[3].split() == "1 10 1 11 5 5 0".split()
    assert oracle.parse_delta_reduce(d)[2] == pytest.approx(0.696969696969697)  # anim.py:343 worked example


def test_anim_matrices_match_cellwise_fill():
    """matrices.anim_matrices vs a literal restatement of pyani_orm.update_comparison_matrices (pyani_orm.py:618-666)."""
    from pyani_b200 import matrices

    tuples, legacy = _gold()
    names = ["NC_014100", "NC_002696", "NC_011916", "NC_010338"]  # deliberately unsorted ids
    lengths = [legacy["org_lengths"][n] for n in names]
    pairs, res = [], []
    for i, a in enumerate(names):
        for j, b in enumerate(names):
            if a == b:
                continue
            ral, qal, ident, errs = tuples[f"{a}_vs_{b}.filter"]
            r = np.zeros(1, dtype=ANIM_DTYPE)[0]
            # integer sums consistent with the golden identity are not recoverable from the tuple; use exact stand-ins
            r["ref_aln_len"], r["qry_aln_len"], r["sim_errs"] = ral, qal, errs
            r["sum_aligned"], r["sum_weighted"] = 10**9, round(ident * 10**9)
            pairs.append((i, j)); res.append(r)
    res = np.array(res, dtype=ANIM_DTYPE)
    got = matrices.anim_matrices(names, lengths, np.array(pairs), res)
    ids = sorted(names)
    # (np.fill_diagonal(df.values, ...) as written in the reference needs a writable view, i.e. pandas < 3)
    want = {k: pd.DataFrame(index=ids, columns=ids, dtype=float) for k in got}
    for k, v in (("identity", 1.0), ("coverage", 1.0), ("sim_errors", 0.0), ("hadamard", 1.0)):
        for n in ids:
            want[k].loc[n, n] = v
    for n, l in zip(names, lengths):
        want["aln_lengths"].loc[n, n] = l
    for (i, j), r in zip(pairs, res):
        q, s = names[i], names[j]
        ident = r["sum_weighted"] / r["sum_aligned"]
        cov = r["ref_aln_len"] / lengths[i]
        want["identity"].loc[q, s] = ident
        want["coverage"].loc[q, s] = cov
        want["aln_lengths"].loc[q, s] = r["ref_aln_len"]
        want["sim_errors"].loc[q, s] = r["sim_errs"]
        want["hadamard"].loc[q, s] = ident * cov
    for k in got:
        assert list(got[k].index) == ids
        assert np.array_equal(got[k].values, want[k].values, equal_nan=True), k


def test_parallel_fasta_ingest(tmp_path):
    """f3: threaded ingest returns, per file, what the reference derives in three Biopython passes + one md5 pass."""
    import hashlib

    from pyani_b200 import fasta

    files = []
    rng = np.random.default_rng(5)
    for i in range(6):
        recs = [rng.choice(list("ACGTN"), size=int(rng.integers(50, 400))) for _ in range(int(rng.integers(1, 4)))]
        p = tmp_path / f"g{i}.fna"
        with open(p, "w") as fh:
            for j, r in enumerate(recs):
                s = "".join(r)
                fh.write(f">rec{j} some description\n" + "\n".join(s[k:k + 60] for k in range(0, len(s), 60)) + "\n")
        files.append((p, recs))
    out = fasta.ingest([p for p, _ in files], threads=4)
    assert [o["stem"] for o in out] == [p.stem for p, _ in files]  # input order kept
    for o, (p, recs) in zip(out, files):
        assert o["md5"] == hashlib.md5(p.read_bytes()).hexdigest()
        assert o["length"] == sum(len(r) for r in recs) and o["ids"] == [f"rec{j}" for j in range(len(recs))]
        assert bytes(o["residues"]).decode() == "".join("".join(r) for r in recs)
        a, off, ids = fasta.read_fasta(p)
        assert (a == o["residues"]).all() and (off == o["offsets"]).all()


def test_comparison_rows_follow_update_comparison_results():
    """f2: Comparison columns exactly as subcmd_anim.update_comparison_results derives them (subcmd_anim.py:438-455)."""
    from pyani_b200 import matrices

    res = np.zeros(3, dtype=ANIM_DTYPE)
    res["ref_aln_len"], res["qry_aln_len"] = [900, 0, 500], [950, 0, 400]
    res["sum_aligned"], res["sum_weighted"], res["sim_errs"] = [2000, 0, 1000], [1800, 0, 990], [100, 0, 5]
    rows = matrices.comparison_rows([11, 22, 33], [1000, 2000, 4000], np.array([[0, 1], [1, 2], [2, 0]]), res, "b200ani 0.1.0", maxmatch=True)
    assert len(rows) == 2  # the alignment-free pair is left out
    a, b = rows
    assert (a["query_id"], a["subject_id"], a["aln_length"], a["sim_errs"]) == (11, 22, 900, 100)
    assert a["identity"] == 0.9 and a["cov_query"] == 0.9 and a["cov_subject"] == 950 / 2000
    assert (b["query_id"], b["subject_id"]) == (33, 11) and b["cov_query"] == 500 / 4000 and b["cov_subject"] == 400 / 1000
    assert a["program"] == "b200ani" and a["version"] == "b200ani 0.1.0" and a["maxmatch"] is True and a["fragsize"] is None


def test_fasta_gt_inside_description():
    """Only a line-initial '>' starts a record (Bio.SeqIO's model, pyani_files.py:141): a '>' in the description is text."""
    from pyani_b200.fasta import _parse

    arr, off, ids = _parse(b">s1 desc a->b\nACGT\nAC\n>s2\nGG\n")
    assert ids == ["s1", "s2"] and off.tolist() == [0, 6, 8] and bytes(arr) == b"ACGTACGG"
    arr, off, ids = _parse(b"\n>only\nAC\r\nGT\r\n")
    assert ids == ["only"] and bytes(arr) == b"ACGT"


def test_fill_order_is_path_order_not_string_order():
    """process_deltadir visits sorted(Path) (anim.py:438): component-wise, so 'a/...' < 'a-b/...' although '-' < '/'.
    The later file wins the symmetric cells."""
    from pathlib import PurePosixPath

    names = ["a", "a-b"]
    files = sorted(PurePosixPath(f"{q}/{q}_vs_{s}.filter") for q in names for s in names if q != s)
    assert [f.parts[0] for f in files] == ["a", "a-b"]  # string order of the joined path would be the reverse
    tuples = {("a", "a-b"): (10, 20, 0.9, 1), ("a-b", "a"): (30, 40, 0.8, 2)}
    res = banim.fill_aniresults({"a": 100, "a-b": 100}, tuples)
    # last visited = a-b_vs_a: its q_len lands in [a-b, a], its s_len in [a, a-b]
    assert res.alignment_lengths.loc["a", "a-b"] == 40 and res.alignment_lengths.loc["a-b", "a"] == 30
    assert res.similarity_errors.loc["a", "a-b"] == 2


def test_aniresults_data_property():
    """The legacy driver iterates ANIResults.data to write its tables (average_nucleotide_identity.py:789)."""
    from pyani_b200.results import _ArrayResults

    for cls in {banim.ANIResults, _ArrayResults}:
        r = cls(["x", "y"], "ANIm")
        r.add_pid("x", "y", 0.5)
        r.add_coverage("x", "y", 0.5, 0.25)
        stems = [stem for _, stem in r.data]
        assert stems == ["ANIm_alignment_lengths", "ANIm_percentage_identity", "ANIm_alignment_coverage", "ANIm_similarity_errors",
                         "ANIm_hadamard"]
        had = dict((stem, df) for df, stem in r.data)["ANIm_hadamard"]
        assert had.loc["x", "y"] == 0.25 and had.loc["y", "x"] == 0.25


def test_native_delta_writer_equals_python_twin(tmp_path):
    """b200ani_write_delta (host-only entry point of the C ABI) writes the same bytes as write_minimal_delta, and the reference's
    record grouping: one '>' line per record pair, alignments by reference start."""
    import ctypes as C

    from pyani_b200 import _lib

    lib = _lib.load()
    rows = np.array([[1, 0, 40, 90, 300, 250, 3], [0, 0, 15, 20, 30, 25, 0], [0, 0, 1, 10, 1, 11, 5], [1, 1, 7, 9, 2, 4, 0]], dtype=np.int64)
    rid, qid = ["r.a", "r.b"], ["q1", "q2"]
    rl, ql = np.array([100, 200], np.int64), np.array([400, 50], np.int64)
    p1, p2 = tmp_path / "n.delta", tmp_path / "p.delta"
    ra = (C.c_char_p * 2)(*[x.encode() for x in rid])
    qa = (C.c_char_p * 2)(*[x.encode() for x in qid])
    rc = lib.b200ani_write_delta(str(p1).encode(), b"/x/R.fna", b"/x/Q.fna", ra, rl.ctypes.data, 2, qa, ql.ctypes.data, 2, rows.ctypes.data, len(rows))
    assert rc == 0
    banim.write_minimal_delta(p2, "/x/R.fna", "/x/Q.fna", rows, rid, qid, rl, ql)
    assert p1.read_bytes() == p2.read_bytes()
    assert p1.read_text().splitlines()[2:5] == [">r.a q1 100 400", "1 10 1 11 5 5 0", "0"]
    bad = rows.copy(); bad[0, 0] = 2
    assert lib.b200ani_write_delta(str(p1).encode(), b"R", b"Q", ra, rl.ctypes.data, 2, qa, ql.ctypes.data, 2, bad.ctypes.data, len(bad)) == -2
