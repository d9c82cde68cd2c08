#!/usr/bin/env python
"""Oracle aligner vs every nucmer output the reference checks in (analysis tool; tests pin the numbers).
  Caulobacter (12 directed pairs, .delta + .filter), C. blochmannia (15 pairs, nucmer 3.1, legacy tarball), JSpecies Group_2
  (2 pairs, .delta only) and the JSpecies concordance table (3 genomes, no deltas)."""
import json
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import oracle  # noqa: E402

FIX = ROOT / "oracle" / "_ref" / "fixtures"


def load(paths):
    out = {}
    for f in paths:
        a, off, _ = oracle.read_fasta(f)
        out[f.stem if not f.name.endswith(".fna.fna") else f.name] = oracle.OGenome(a, off)
    return out


def compare(R, Q, delta, filt=None):
    rows, st = oracle.anim_align(R, Q)
    nd = oracle.parse_delta_text(delta)
    o = {}
    _, _, n_idy, _ = oracle.parse_delta_reduce(nd[:, :7])
    n_ral = oracle.parse_delta_reduce(nd[:, :7])[0]
    ral, qal, idy, errs = oracle.parse_delta_reduce(rows[:, :7]) if len(rows) else (0, 0, 0, 0)
    o["d_idy_unf"] = 100 * (idy - n_idy)
    o["len_ratio_unf"] = ral / n_ral
    ours, nuc = set(map(tuple, rows[:, 2:7].tolist())), set(map(tuple, nd[:, 2:7].tolist()))
    o["rows"] = (len(ours & nuc), len(nuc), len(ours))
    if filt is not None:
        nf = oracle.parse_delta_text(filt)
        f_ral, f_qal, f_idy, _ = oracle.parse_delta_reduce(nf[:, :7])
        keep = oracle.filter_1to1(rows)
        k_ral, k_qal, k_idy, _ = oracle.parse_delta_reduce(rows[keep][:, :7])
        o["d_idy_filt"] = 100 * (k_idy - f_idy)
        o["len_ratio_filt"] = k_ral / f_ral
        o["n_filt"] = (int(keep.sum()), len(nf))
    return o


def main():
    which = sys.argv[1:] or ["caulo", "bloch", "group2", "conc"]
    jobs = []
    if "caulo" in which:
        g = load(sorted((FIX / "sequences").glob("*.fna")))
        for f in sorted((FIX / "anim_deltadir").glob("*.delta")):
            r, q = f.stem.split("_vs_")
            jobs.append(("caulo " + f.stem, g[r], g[q], f, f.with_suffix(".filter")))
    if "bloch" in which:
        g = load(sorted((FIX / "C_blochmannia").glob("*.fna")))
        for f in sorted((FIX / "C_blochmannia_nucmer").glob("*.delta")):
            r, q = f.stem.split("_vs_")
            jobs.append(("bloch " + r[4:13] + "_" + q[4:13], g[r], g[q], f, f.with_suffix(".filter")))
    if "group2" in which:
        d = FIX / "jspecies_group2"
        g = load(sorted(d.glob("*.fna")))
        for r, q in (("TS32_31", "TS43_20"), ("TS43_20", "TS32_31")):
            jobs.append((f"group2 {r}_{q}", g[r], g[q], d / f"{r}.fna_vs_{q}.fna.delta", None))
    with ThreadPoolExecutor(8) as ex:
        res = list(ex.map(lambda j: compare(*j[1:]), jobs))
    tot_same = tot_nuc = 0
    for j, o in zip(jobs, res):
        same, nn, no = o["rows"]
        tot_same += same; tot_nuc += nn
        print(f"{j[0]:48s} unf d_idy {o['d_idy_unf']:+.3f} len {o['len_ratio_unf']:.4f} rows {same}/{nn} (ours {no})" +
              (f" | filt d_idy {o['d_idy_filt']:+.3f} len {o['len_ratio_filt']:.4f} n {o['n_filt']}" if "d_idy_filt" in o else ""))
    if tot_nuc:
        print(f"rows reproduced {tot_same}/{tot_nuc} = {tot_same / tot_nuc:.3f}")
    if "conc" in which:
        tab = json.load(open(ROOT / "tests/golden/jspecies_concordance.json"))["ANIm"]
        g = load(sorted((FIX / "concordance").glob("*.fna")))
        names = [n[:-4] for n in tab["names"]]
        for i, r in enumerate(names):
            for k, q in enumerate(names):
                if i == k:
                    continue
                res, kept, st = oracle.anim_pair(g[r], g[q])
                unf = oracle.anim_pair(g[r], g[q], nofilter=True)[0]
                idy = 100 * res["sum_weighted"] / res["sum_aligned"]
                idu = 100 * unf["sum_weighted"] / unf["sum_aligned"]
                print(f"conc {r[4:13]} {q[4:13]} ours {idy:.3f} (unfiltered {idu:.3f}) jspecies {tab['values'][i][k]} / {tab['values'][k][i]}  dev {idy - tab['values'][i][k]:+.3f}"
                      f"  aln {res['ref_aln_len']} n {res['n_alignments']}")


if __name__ == "__main__":
    main()
