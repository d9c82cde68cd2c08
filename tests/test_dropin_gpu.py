"""The job-graph backends driven by the UNMODIFIED reference on a real GPU.

The reference (pyani v0.3.0-alpha, pip-installed into baseline/_ref by the recipe in DESIGN.md; Biopython / intervaltree /
matplotlib stubbed by oracle/refstubs) builds the job graph (`anim.generate_nucmer_jobs`, `anib.make_job_graph`), this package's
`run_dependency_graph` runs it on the engine and writes the files the external aligners would have written, and the
reference's own `anim.process_deltadir` / `anib.process_blast` parse them.  The result must equal `calculate_anim` /
`calculate_anib` (same engine, no files) cell for cell."""
import sys
import time
from pathlib import Path

import numpy as np
import pytest

from conftest import ROOT, fixture_genomes

pytestmark = pytest.mark.gpu
REF = ROOT / "baseline" / "_ref"


@pytest.fixture(scope="module")
def ref():
    if not (REF / "pyani" / "anim.py").exists():
        pytest.skip("reference not installed under baseline/_ref (see DESIGN.md)")
    sys.path[:0] = [str(ROOT / "oracle" / "refstubs"), str(REF)]
    from pyani import anib, anim, pyani_files, pyani_jobs  # noqa: F401  (the unmodified reference)

    return {"anim": anim, "anib": anib, "files": pyani_files}


def _small_fastas(tmp_path, n=4, length=150_000):
    from pyani_b200 import synth

    gs = synth.make_set(2, 2, length, seed=77)[:n]
    paths = []
    for g in gs:
        p = tmp_path / f"{g.name}.fna"
        with open(p, "wb") as fh:
            for c in range(len(g.contig_off) - 1):
                seq = bytes(g.residues[int(g.contig_off[c]):int(g.contig_off[c + 1])])
                fh.write(b">%s_rec%d some description\n" % (g.name.encode(), c))
                for k in range(0, len(seq), 80):
                    fh.write(seq[k:k + 80] + b"\n")
        paths.append(p)
    return paths


def test_anim_job_graph_through_the_reference(engine, ref, tmp_path):
    from pyani_b200 import anim as banim

    paths = _small_fastas(tmp_path)
    outdir = tmp_path / "out"
    outdir.mkdir()
    jobs = ref["anim"].generate_nucmer_jobs(paths, outdir, Path("nucmer"), Path("delta-filter"), False, "T")
    assert len(jobs) == len(paths) * (len(paths) - 1)
    assert banim.run_dependency_graph(jobs, workers=8, engine=engine) == 0
    org_lengths = ref["files"].get_sequence_lengths(paths)
    got = ref["anim"].process_deltadir(outdir / "nucmer_output", org_lengths)  # the reference parses what the engine wrote
    want = banim.calculate_anim(paths, org_lengths, engine=engine)
    for tab in ("alignment_lengths", "similarity_errors", "percentage_identity", "alignment_coverage"):
        g, w = getattr(got, tab), getattr(want, tab).loc[getattr(got, tab).index, getattr(got, tab).columns]
        assert np.array_equal(g.values.astype(float), w.values.astype(float), equal_nan=True), tab
    # every .filter is a subset of its .delta, and parse_delta of the reference reads both
    for f in sorted((outdir / "nucmer_output").glob("*/*.filter")):
        d = f.with_suffix(".delta")
        assert ref["anim"].parse_delta(f)[0] <= ref["anim"].parse_delta(d)[0]


def test_anib_job_graph_through_the_reference(engine, ref, tmp_path):
    from pyani_b200 import anib as banib

    paths = _small_fastas(tmp_path, n=3)
    outdir = tmp_path / "blastn_output"
    outdir.mkdir()
    fragfiles, fraglengths = ref["anib"].fragment_fasta_files(paths, outdir, 1020)
    jobgraph = ref["anib"].make_job_graph(paths, fragfiles, ref["anib"].make_blastcmd_builder("ANIb", outdir))
    assert banib.run_dependency_graph(jobgraph, engine=engine) == 0
    org_lengths = ref["files"].get_sequence_lengths(paths)
    got = ref["anib"].process_blast(outdir, org_lengths, fraglengths, mode="ANIb")
    want = banib.calculate_anib(paths, org_lengths, engine=engine)
    for tab in ("alignment_lengths", "similarity_errors", "alignment_coverage"):
        g, w = getattr(got, tab), getattr(want, tab).loc[getattr(got, tab).index, getattr(got, tab).columns]
        assert np.array_equal(g.values.astype(float), w.values.astype(float), equal_nan=True), tab
    # identity goes through BLAST's 3-decimal pident column in the file: equal to 1e-9
    g, w = got.percentage_identity, want.percentage_identity.loc[got.percentage_identity.index, got.percentage_identity.columns]
    assert np.allclose(g.values.astype(float), w.values.astype(float), atol=1e-9)


def test_batched_backend_cost(engine, ref, tmp_path):
    """The whole job graph costs one batched device pass plus file writing: not more than a few times the in-memory call."""
    from pyani_b200 import anim as banim

    paths = _small_fastas(tmp_path, n=4, length=300_000)
    outdir = tmp_path / "out"
    outdir.mkdir()
    jobs = ref["anim"].generate_nucmer_jobs(paths, outdir, Path("nucmer"), Path("delta-filter"), False, "T")
    banim.calculate_anim(paths, engine=engine)  # warm
    t0 = time.perf_counter(); banim.calculate_anim(paths, engine=engine); t_mem = time.perf_counter() - t0
    t0 = time.perf_counter(); assert banim.run_dependency_graph(jobs, engine=engine) == 0; t_files = time.perf_counter() - t0
    assert t_files < 3 * t_mem + 0.5, (t_files, t_mem)
