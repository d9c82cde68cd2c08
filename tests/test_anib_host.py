"""Host-side mirror of the reference's ANIb interfaces (pyani_b200/anib.py) -- no GPU needed."""
import json

import numpy as np
import pytest
import pandas as pd

import oracle
from conftest import GOLDEN, fixture_genomes
from pyani_b200 import anib as banib
from pyani_b200.engine import ANIB_DTYPE


def test_fill_aniresults_matches_reference_process_blast():
    """anib.process_blast run by the reference on its own BLAST+ fixtures vs fill_aniresults fed the per-file
    tuples: every cell of all four tables (incl. the symmetric, last-file-wins similarity errors)."""
    tuples = json.load(open(GOLDEN / "anib_parse_blast_tab.json"))
    legacy = json.load(open(GOLDEN / "anib_process_blast.json"))
    per_file = {tuple(k.split("_vs_")): tuple(v) for k, v in tuples.items()}
    res = banib.fill_aniresults(legacy["org_lengths"], per_file)
    for tab in ("alignment_lengths", "similarity_errors", "percentage_identity", "alignment_coverage"):
        g = legacy[tab]
        want = pd.DataFrame(g["values"], index=g["index"], columns=g["columns"])
        got = getattr(res, tab).loc[g["index"], g["columns"]]
        assert (got.values == want.values).all(), tab


def test_parse_result_and_fragment_lengths():
    r = np.zeros(1, dtype=ANIB_DTYPE)[0]
    r["aln_length"], r["sim_errs"], r["pid_milli_sum"], r["n_pass"] = 1818, 42, 98750 + 97000, 2
    assert banib.parse_result(r) == (1818, 42, (98.75 + 97.0) / 2)
    assert banib.parse_result(np.zeros(1, dtype=ANIB_DTYPE)[0]) == (0, 0, 0.0)  # anib.py:661-663
    fl = banib.fragment_lengths([2500, 1020, 5], 1020)  # anib.py:192-200: ids run across records, short tails kept
    assert fl == {"frag00001": 1020, "frag00002": 1020, "frag00003": 460, "frag00004": 1020, "frag00005": 5}


def test_fragment_counts_of_fixture_genomes():
    files = fixture_genomes()
    if not files:
        import pytest

        pytest.skip("fixture genomes not copied")
    d = banib.get_fraglength_dict(files)
    # tests/fixtures/anib/fragfiles hold 3,940 / 5,371 / 3,964 / 4,565 fragments
    assert [len(d[f.stem]) for f in files] == [3940, 5371, 3964, 4565]
    assert max(max(v.values()) for v in d.values()) == 1020  # tests/test_anib.py:317-324


def test_minimal_blast_tab_round_trip(tmp_path):
    """f1: a .blast_tab written from the engine's per-fragment rows reduces (anib.parse_blast_tab restatement, pinned to the
    reference's fixtures in test_oracle_anib_reduce) to the sums the engine reports for the pair."""
    from pyani_b200 import anib as banib

    rng = np.random.default_rng(3)
    n = 40
    qlen = np.full(n, 1020)
    gaps = rng.integers(0, 30, n)
    mism = rng.integers(0, 150, n)
    nident = rng.integers(800, 1000, n)
    length = nident + mism + gaps
    rows = np.stack([np.arange(n) * 2, length, mism, gaps, nident, qlen, 2 * nident, (100000 * nident + length // 2) // length], axis=1)
    f = tmp_path / "q_vs_s.blast_tab"
    banib.write_minimal_blast_tab(f, "subject", rows, slen=4_000_000)
    parsed = []
    for line in open(f):
        c = line.rstrip("\n").split("\t")
        assert len(c) == 15 and c[0].startswith("frag") and c[1] == "subject"
        parsed.append((int(c[0][4:]), int(c[2]), int(c[3]), float(c[4]), int(c[6]), int(c[14])))
    aln, errs, pid = oracle.parse_blast_tab_reduce(np.asarray(parsed, dtype=np.float64))
    keep = ((length - gaps) / qlen > 0.7) & ((length - gaps - mism) / qlen > 0.3)
    assert aln == int((length - gaps)[keep].sum()) and errs == int((mism + gaps)[keep].sum())
    assert abs(pid - np.mean(np.round(rows[keep, 7] / 1000, 3))) < 1e-9


class _Job:
    def __init__(self, name, command):
        self.name, self.command, self.dependencies = name, command, []

    def add_dependency(self, job):
        self.dependencies.append(job)


class _FakeEngine:
    def __init__(self):
        self.names, self.calls = [], []

    def add_genome(self, arr, off, name, ids=None):
        self.names.append((name, int(off[-1])))
        return len(self.names) - 1

    def genome_length(self, gid):
        return self.names[gid][1]

    def anib_rows(self, pairs, fragsize=1020):
        per_pair = np.array([[0, 1000, 50, 10, 940, 1020, 1700, 94000], [1, 400, 10, 0, 390, 412, 750, 97500]])
        self.batches = getattr(self, "batches", 0) + 1
        for q, s in np.asarray(pairs).tolist():
            self.calls.append((self.names[q][0], self.names[s][0], fragsize))
        m = len(pairs)
        return None, np.concatenate([per_pair] * m), np.arange(m + 1) * 2


def test_anib_job_graph_backend_writes_blast_tab(tmp_path):
    """f1: the ANIb job graph of anib.make_job_graph (anib.py:296-347), command lines as construct_makeblastdb_cmd
    (anib.py:378-391) and construct_blastn_cmdline (anib.py:451-471) build them."""
    fa, fb = tmp_path / "A.fna", tmp_path / "B.fna"
    fa.write_text(">a1 d\nACGTACGTAC\nGTACGTAC\n")
    fb.write_text(">b1 d\nACGTACGTACGTAC\n")
    out = tmp_path / "blastn_output"
    fmt = "'6 qseqid sseqid length mismatch pident nident qlen slen qstart qend sstart send positive ppos gaps'"
    dbs = {n: _Job(f"db_{n}", f"makeblastdb -dbtype nucl -in {f} -title {n} -out {out / f.name}") for n, f in (("A", fa), ("B", fb))}
    jobs = []
    for q, s in (("A", "B"), ("B", "A")):
        j = _Job(f"exe_{q}{s}", f"blastn -out {out}/{q}_vs_{s}.blast_tab -query {out}/{q}-fragments.fna -db {out}/{s}.fna "
                                f"-xdrop_gap_final 150 -dust no -evalue 1e-15 -max_target_seqs 1 -outfmt {fmt} -task blastn")
        j.add_dependency(dbs[s])
        jobs.append(j)
    eng = _FakeEngine()
    assert banib.run_dependency_graph(jobs, workers=8, engine=eng) == 0
    assert sorted(eng.calls) == [("A", "B", 1020), ("B", "A", 1020)] and eng.batches == 1  # ONE batched pass for the whole graph
    lines = open(out / "A_vs_B.blast_tab").read().splitlines()
    assert len(lines) == 2 and lines[0].split("\t") == ["frag00001", "B", "1000", "50", "94.000", "940", "1020", "14", "0", "0", "0", "0", "940", "94.000", "10"]
    rows = [(int(c[0][4:]), int(c[2]), int(c[3]), float(c[4]), int(c[6]), int(c[14])) for c in (l.split("\t") for l in lines)]
    assert oracle.parse_blast_tab_reduce(np.asarray(rows, dtype=np.float64)) == (990 + 400, 60 + 10, pytest.approx((94.0 + 97.5) / 2))
