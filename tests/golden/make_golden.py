#!/usr/bin/env python
"""Generate the committed golden vectors by running THE REFERENCE ITSELF in the build container.

Run from the repo root:  python tests/golden/make_golden.py
Needs /root/reference (read-only) and the stub modules under oracle/refstubs (Biopython /
intervaltree / matplotlib are not installed here).  Outputs (small, committed):
  tests/golden/tetra_reference.json      Z-scores + correlations from pyani.tetra on the 4 fixture
                                         genomes and on small inline edge-case FASTAs
  tests/golden/anim_parse_delta.json     anim.parse_delta tuples on every checked-in .delta/.filter
  tests/golden/anim_alignments.npz       the 7-int alignment headers of those files (inputs)
It also copies the fixture genomes the GPU tests read into oracle/_ref/fixtures/ (git-ignored, but
shipped to the GPU box by gpurun, where /root/reference does not exist).
"""
import json
import shutil
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
REF = Path("/root/reference")
sys.path[:0] = [str(ROOT / "oracle" / "refstubs"), str(REF), str(ROOT)]

from pyani import anim, tetra  # noqa: E402  (the reference)
import oracle  # noqa: E402

OUT = ROOT / "tests" / "golden"
FIX = ROOT / "oracle" / "_ref" / "fixtures"

EDGE_FASTAS = {
    # name -> FASTA text; exercise lower case, Ns/IUPAC, multi-record, short records (L<5) etc.
    "edge_multi": ">r1 first\nACGTTGCAAGGCTTAACCGGTTAAGGCCTTAGCATCGATCGGATCGATTTAGC\nGATTACAGATTACAGGGCATCGA\n>r2\nTTGACCAGTAGGCATTTACGAGCAGT\n>r3\nACG\n>r4\nACGT\n>r5\nACGTA\n",
    "edge_lower_n": ">x\nacgtnnacgtacgatcgatcgatcgRYacgatgcatgcatgcatcgatcgatgcatgcatgcNacgt\nACGATCGATCAGCTAGCTAGCATCGATCGATCGATCGACTAGCTAGCATCGATGCAT\n",
}


def rand_fasta(seed, n_rec, length, p_n=0.002):
    rng = np.random.default_rng(seed)
    out = []
    for r in range(n_rec):
        s = rng.choice(list("ACGT"), size=length, p=[0.2, 0.3, 0.3, 0.2])
        s[rng.random(length) < p_n] = "N"
        s = "".join(s)
        out.append(">rec%d d\n" % r + "\n".join(s[i : i + 70] for i in range(0, len(s), 70)) + "\n")
    return "".join(out)


def jspecies_tables(path):
    """tests/fixtures/concordance/jspecies_output.tab -> {"Tetra" | "ANIb" | "ANIm": {"names": [...], "values": [[...]]}}
    (row = first column label, column = header label; '---' on the diagonal -> None).  The targets of the reference's
    own concordance tests (tests/test_concordance.py:169-302)."""
    out, cur, names = {}, None, None
    for raw in list(open(path)) + [""]:
        line = raw.strip()
        if line in ("Tetra", "ANIb", "ANIm"):
            cur, names = line, None
            out[cur] = {"names": [], "rows": [], "values": []}
        elif cur and names is None and line:
            names = [c for c in line.split("\t") if c]
            out[cur]["names"] = names
        elif cur and line:
            cells = [c for c in line.split("\t") if c]
            out[cur]["rows"].append(cells[0])
            out[cur]["values"].append([None if c == "---" else float(c) for c in cells[1:]])
        elif cur and not line and names is not None:
            cur = None
    return out


def main():
    OUT.mkdir(parents=True, exist_ok=True)
    FIX.mkdir(parents=True, exist_ok=True)
    json.dump(jspecies_tables(REF / "tests/fixtures/concordance/jspecies_output.tab"), open(OUT / "jspecies_concordance.json", "w"), indent=1)
    # JSpecies 1.2.1 on the four Caulobacter fixture genomes (tests/test_JSpecies/README.md) and the nucmer deltas JSpecies
    # wrote for an independent pair of draft genomes (Group_2): never used to calibrate anything here
    json.dump(jspecies_tables(REF / "tests/test_JSpecies/jspecies_results.tab"), open(OUT / "jspecies_caulobacter.json", "w"), indent=1)
    g2 = FIX / "jspecies_group2"
    g2.mkdir(exist_ok=True)
    for f in sorted((REF / "tests/test_JSpecies/Group_2").glob("TS*")):
        if f.suffix in (".fna", ".delta"):
            shutil.copyfile(f, g2 / f.name)
    EDGE_FASTAS["edge_rand3x4k"] = rand_fasta(11, 3, 4000)
    EDGE_FASTAS["edge_rand1x20k"] = rand_fasta(12, 1, 20000, p_n=0.0)

    # ---- fixture genomes for the GPU box
    for sub in ("fixtures/sequences", "test_targets/legacy_scripts/C_blochmannia", "fixtures/concordance"):
        dst = FIX / Path(sub).name
        dst.mkdir(exist_ok=True)
        for f in sorted((REF / "tests" / sub).glob("*.fna")):
            shutil.copyfile(f, dst / f.name)
    dd = FIX / "anim_deltadir"
    dd.mkdir(exist_ok=True)
    for f in sorted((REF / "tests/fixtures/anim/deltadir").glob("*/*_vs_*")):
        if f.suffix in (".delta", ".filter"):
            shutil.copyfile(f, dd / f.name)

    # nucmer 3.1 output of the legacy driver for the six C. blochmannia genomes (15 pairs x {.delta, .filter}): pins the aligner
    # on genomes no constant was fitted to
    import tarfile

    bn = FIX / "C_blochmannia_nucmer"
    bn.mkdir(exist_ok=True)
    with tarfile.open(REF / "tests/test_targets/legacy_scripts/ANIm_mpl_Linux_3.1/nucmer_output.tar.gz") as tf:
        for m in tf.getmembers():
            if m.isfile() and m.name.endswith((".delta", ".filter")):
                (bn / Path(m.name).name).write_bytes(tf.extractfile(m).read())

    # ---- TETRA
    gold = {"tetramers": oracle.TETRAMERS, "genomes": {}, "edge": {}}
    zs = {}
    for f in sorted((REF / "tests/fixtures/sequences").glob("*.fna")):
        z = tetra.calculate_tetra_zscore(f)
        zs[f.stem] = z
        gold["genomes"][f.stem] = {k: z[k] for k in sorted(z)}
        print("tetra", f.stem, len(z))
    corr = tetra.calculate_correlations(zs)
    gold["corr_py312_sum"] = {"orgs": list(corr.index), "matrix": corr.values.tolist()}
    # the reference's own checked-in tables
    gold["zscore_json_NC_002696"] = json.load(open(REF / "tests/fixtures/targets/tetra/zscore.json"))
    tab = [l.rstrip("\n").split("\t") for l in open(REF / "tests/target_TETRA_output/TETRA_correlations.tab")]
    gold["corr_checked_in"] = {"orgs": tab[0][1:], "matrix": [[float(x) for x in r[1:]] for r in tab[1:]]}
    with tempfile.TemporaryDirectory() as td:
        for name, txt in EDGE_FASTAS.items():
            p = Path(td) / (name + ".fna")
            p.write_text(txt)
            z = tetra.calculate_tetra_zscore(p)
            gold["edge"][name] = {"fasta": txt, "z": {k: z[k] for k in sorted(z)}}
    json.dump(gold, open(OUT / "tetra_reference.json", "w"))

    # ---- ANIm reduction
    files = sorted((REF / "tests/fixtures/anim/deltadir").glob("*/*.filter"))
    files += sorted((REF / "tests/fixtures/anim/deltadir").glob("*/*.delta"))
    files += [REF / "tests/fixtures/anim/test.delta"]
    tuples, arrays = {}, {}
    for f in files:
        tuples[f.name] = list(anim.parse_delta(f))
        arrays[f.name] = oracle.parse_delta_text(f).astype(np.int32)
    json.dump(tuples, open(OUT / "anim_parse_delta.json", "w"), indent=0)
    # legacy matrices: the reference's process_deltadir on its own fixture directory (anim.py:415-497)
    from pyani import pyani_files

    seqdir = REF / "tests/fixtures/anim/sequences"
    org_lengths = pyani_files.get_sequence_lengths(pyani_files.get_fasta_files(seqdir))
    res = anim.process_deltadir(REF / "tests/fixtures/anim/deltadir", org_lengths)
    legacy = {"org_lengths": org_lengths}
    for name in ("alignment_lengths", "similarity_errors", "percentage_identity", "alignment_coverage"):
        df = getattr(res, name)
        legacy[name] = {"index": list(df.index), "columns": list(df.columns), "values": df.values.tolist()}
    legacy["deltadir_result_csv"] = (REF / "tests/fixtures/anim/dataframes/deltadir_result.csv").read_text()
    json.dump(legacy, open(OUT / "anim_process_deltadir.json", "w"))
    np.savez_compressed(OUT / "anim_alignments.npz", **arrays)
    print("anim", len(tuples))

    # ---- ANIb reduction: the reference's parse_blast_tab on its checked-in BLAST+ output.  The function writes a
    # <file>.dataframe next to its input (anib.py:666), so the fixtures are copied out of the read-only tree first.
    from pyani import anib

    blast_gold, blast_rows = {}, {}
    with tempfile.TemporaryDirectory() as td:
        for f in sorted((REF / "tests/fixtures/anib/blastn").glob("*.blast_tab")):
            dst = Path(td) / f.name
            shutil.copyfile(f, dst)
            aln, errs, pid = anib.parse_blast_tab(dst, None, mode="ANIb")
            blast_gold[f.stem] = [int(aln), int(errs), float(pid)]
            # inputs of the restatement: fragment number, length, mismatch, pident, qlen, gaps per HSP row, file order
            rows = []
            for line in open(f):
                c = line.rstrip("\n").split("\t")
                rows.append((int(c[0].replace("frag", "")), int(c[2]), int(c[3]), float(c[4]), int(c[6]), int(c[14])))
            blast_rows[f.stem] = np.asarray(rows, dtype=np.float64)
    with tempfile.TemporaryDirectory() as td:
        for f in sorted((REF / "tests/fixtures/anib/blastn").glob("*.blast_tab")):
            shutil.copyfile(f, Path(td) / f.name)
        seqdir = REF / "tests/fixtures/anib/sequences"
        org_lengths_b = pyani_files.get_sequence_lengths(pyani_files.get_fasta_files(seqdir))
        resb = anib.process_blast(Path(td), org_lengths_b, None, mode="ANIb")
        legacy_b = {"org_lengths": org_lengths_b}
        for name in ("alignment_lengths", "similarity_errors", "percentage_identity", "alignment_coverage"):
            df = getattr(resb, name)
            legacy_b[name] = {"index": list(df.index), "columns": list(df.columns), "values": df.values.tolist()}
        json.dump(legacy_b, open(OUT / "anib_process_blast.json", "w"))
    json.dump(blast_gold, open(OUT / "anib_parse_blast_tab.json", "w"), indent=0)
    np.savez_compressed(OUT / "anib_blast_rows.npz", **blast_rows)
    print("anib", len(blast_gold))


if __name__ == "__main__":
    main()
