"""CPU oracle for the pyani pairwise hot path -- TEST INFRASTRUCTURE, not product code.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package.  ``pyani_b200`` (the product) never does: it fails loudly if its CUDA
library is missing instead of falling back here.

The arithmetic lives in the plain-C files next to this one (each function cites the reference
file:line it restates); this module is the ctypes shim plus the two text parsers
(`.delta` headers, anim.py:366-393) needed to feed the reference's checked-in aligner output to it.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = None


def build(force: bool = False) -> Path:
    so = _HERE / "liboracle.so"
    srcs = sorted(_HERE.glob("oracle_*.c")) + sorted(_HERE.glob("*.h"))
    if force or not so.exists() or any(s.stat().st_mtime > so.stat().st_mtime for s in srcs):
        subprocess.run(["make", "-s", "-C", str(_HERE), "liboracle.so"], check=True)
    return so


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        so = _HERE / "liboracle.so"
        if os.environ.get("ORACLE_SO"):  # parameter-sweep builds (tests/tools only)
            _LIB = ctypes.CDLL(os.environ["ORACLE_SO"])
            return _LIB
        if not so.exists() or os.environ.get("ORACLE_REBUILD"):
            build()
        else:
            try:
                build()
            except Exception:  # GPU box without make/gcc still has the shipped .so
                pass
        _LIB = ctypes.CDLL(str(so))
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


# ----------------------------------------------------------------------------- TETRA
def tetra_counts(ascii_u8: np.ndarray, contig_off: np.ndarray):
    ascii_u8 = np.ascontiguousarray(ascii_u8, dtype=np.uint8)
    contig_off = np.ascontiguousarray(contig_off, dtype=np.uint64)
    c2, c3, c4 = (np.zeros(k, dtype=np.int64) for k in (16, 64, 256))
    lib().orc_tetra_counts(_p(ascii_u8, ctypes.c_uint8), _p(contig_off, ctypes.c_uint64),
                           ctypes.c_uint32(len(contig_off) - 1),
                           _p(c2, ctypes.c_int64), _p(c3, ctypes.c_int64), _p(c4, ctypes.c_int64))
    return c2, c3, c4


def tetra_zscores(c2, c3, c4):
    z = np.zeros(256, dtype=np.float64)
    obs = np.zeros(256, dtype=np.uint8)
    lib().orc_tetra_zscores(_p(c2, ctypes.c_int64), _p(c3, ctypes.c_int64), _p(c4, ctypes.c_int64),
                            _p(z, ctypes.c_double), _p(obs, ctypes.c_uint8))
    return z, obs


def tetra_corr(z: np.ndarray, mask: np.ndarray | None = None):
    z = np.ascontiguousarray(z, dtype=np.float64)
    n = z.shape[0]
    mask = np.ones(256, np.uint8) if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
    corr = np.zeros((n, n), dtype=np.float64)
    lib().orc_tetra_corr(_p(z, ctypes.c_double), _p(mask, ctypes.c_uint8), ctypes.c_int(n), _p(corr, ctypes.c_double))
    return corr


TETRAMERS = ["".join("ACGT"[(t >> s) & 3] for s in (6, 4, 2, 0)) for t in range(256)]


# ----------------------------------------------------------------------------- ANIm reduction
def parse_delta_text(path) -> np.ndarray:
    """Rows (ref_seq_id, qry_seq_id, rs, re, qs, qe, errs) from a .delta/.filter file.

    Restates the line classification of anim.py:366-393: '>' lines switch the current sequence
    pair, 7-field lines are alignment headers, everything else (indel offsets) is ignored.
    Sequence names are mapped to integer ids per side (ids only need to be distinct per name).
    """
    rid, qid, rows = {}, {}, []
    cr = cq = -1
    with open(path) as fh:
        for raw in fh:
            f = raw.strip().split()
            if not f or f[0] == "NUCMER":
                continue
            if f[0].startswith(">"):
                cr = rid.setdefault(f[0][1:], len(rid))
                cq = qid.setdefault(f[1], len(qid))
            elif len(f) == 7:
                rows.append((cr, cq, int(f[0]), int(f[1]), int(f[2]), int(f[3]), int(f[4])))
    return np.asarray(rows, dtype=np.int64).reshape(-1, 7)


def parse_delta_reduce(aln: np.ndarray):
    """-> (ref_aln_len, qry_aln_len, identity, sim_errs) exactly as anim.parse_delta (anim.py:411)."""
    aln = np.ascontiguousarray(aln, dtype=np.int64).reshape(-1, 7)
    out = np.zeros(5, dtype=np.int64)
    lib().orc_parse_delta_reduce(_p(aln, ctypes.c_int64), ctypes.c_int64(aln.shape[0]), _p(out, ctypes.c_int64))
    if out[2] == 0:
        raise ZeroDivisionError("no alignments (anim.py:396 divides by sum(aligned_bases))")
    return int(out[0]), int(out[1]), int(out[3]) / int(out[2]), int(out[4])


# ----------------------------------------------------------------------------- FASTA (host helper)
def read_fasta(path):
    """-> (ascii uint8 array of all residues, contig_off uint64[n+1], [record ids]). Ns are kept."""
    ids, chunks, lens = [], [], []
    cur = None
    with open(path, "rb") as fh:
        for line in fh:
            if line.startswith(b">"):
                if cur is not None:
                    lens.append(sum(len(c) for c in cur)); chunks.extend(cur)
                ids.append(line[1:].split()[0].decode() if line[1:].split() else "")
                cur = []
            elif cur is not None:
                s = line.strip()
                if s:
                    cur.append(s)
    if cur is not None:
        lens.append(sum(len(c) for c in cur)); chunks.extend(cur)
    data = np.frombuffer(b"".join(chunks), dtype=np.uint8).copy()
    off = np.zeros(len(lens) + 1, dtype=np.uint64)
    off[1:] = np.cumsum(np.asarray(lens, dtype=np.uint64))
    return data, off, ids


# ----------------------------------------------------------------------------- ANIm aligner (oracle_anim.c)
ANIM_NSTATS = 16
ANIM_STAT_NAMES = ["anchors", "links", "cut_links", "segments", "gap_cells", "ext_cells", "ext_tiles", "seed_hits", "raw_anchors", "bridges",
                   "bridge_cells", "word_hits", "word_anchors", "", "", ""]


class OGenome:
    """Oracle-side genome handle (byte-per-base g-space + minimizer index: k-mers with <= 16 occurrences, repeats flagged)."""

    def __init__(self, ascii_u8: np.ndarray, contig_off: np.ndarray, anim_index: bool = True):
        """anim_index=False: skip the ANIm seed index (callers that only run ANIb; the ANIb index is built on first use)."""
        L = lib()
        L.orc_genome_create.restype = ctypes.c_void_p
        L.orc_genome_free.argtypes = [ctypes.c_void_p]
        L.orc_genome_n_mini.restype = ctypes.c_int64
        L.orc_genome_n_mini.argtypes = [ctypes.c_void_p]
        a = np.ascontiguousarray(ascii_u8, dtype=np.uint8)
        o = np.ascontiguousarray(contig_off, dtype=np.uint64)
        self.n_bases = int(o[-1])
        self.contig_off = o
        self._args = (a, o)
        self._h = ctypes.c_void_p(L.orc_genome_create(_p(a, ctypes.c_uint8), _p(o, ctypes.c_uint64), ctypes.c_uint32(len(o) - 1))) if anim_index else None
        self._hb = None  # ANIb index (shorter seeds), built on first use

    @property
    def anib_handle(self):
        if self._hb is None:
            L = lib()
            L.orc_genome_create_anib.restype = ctypes.c_void_p
            a, o = self._args
            self._hb = ctypes.c_void_p(L.orc_genome_create_anib(_p(a, ctypes.c_uint8), _p(o, ctypes.c_uint64), ctypes.c_uint32(len(o) - 1)))
        return self._hb

    def __del__(self):
        try:
            if self._h:
                lib().orc_genome_free(self._h)
                self._h = None
            if self._hb:
                lib().orc_genome_free(self._hb)
                self._hb = None
        except Exception:
            pass

    @property
    def n_mini(self):
        return int(lib().orc_genome_n_mini(self._h))

    def minimizers(self):
        n = self.n_mini
        k = np.zeros(n, np.uint32); pf = np.zeros(n, np.uint32)
        lib().orc_genome_minimizers.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        lib().orc_genome_minimizers(self._h, k.ctypes.data, pf.ctypes.data)
        return k, pf


def anim_align(R: OGenome, Q: OGenome, maxmatch: bool = False, cap: int = 1 << 16):
    """Unfiltered alignments of R (nucmer reference) vs Q: rows (rc, qc, rs, re, qs, qe, errs, score) + stats dict."""
    L = lib()
    L.orc_anim_align.restype = ctypes.c_int64
    L.orc_anim_align.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]
    rows = np.zeros((cap, 8), dtype=np.int64)
    stats = np.zeros(ANIM_NSTATS, dtype=np.int64)
    n = L.orc_anim_align(R._h, Q._h, int(bool(maxmatch)), rows.ctypes.data, cap, stats.ctypes.data)
    assert n <= cap, "alignment capacity exceeded"
    return rows[:n].copy(), {k: v for k, v in zip(ANIM_STAT_NAMES, stats.tolist()) if k}


def filter_1to1(rows: np.ndarray) -> np.ndarray:
    """delta-filter -1 restatement: boolean keep mask for alignment rows (n x 8, score column unused)."""
    rows = np.ascontiguousarray(rows, dtype=np.int64)
    if rows.shape[1] == 7:
        rows = np.concatenate([rows, np.zeros((len(rows), 1), np.int64)], axis=1)
    keep = np.zeros(len(rows), dtype=np.uint8)
    L = lib()
    L.orc_filter_1to1.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]
    L.orc_filter_1to1(rows.ctypes.data, len(rows), keep.ctypes.data)
    return keep.astype(bool)


def anim_pair(R: OGenome, Q: OGenome, nofilter: bool = False, maxmatch: bool = False):
    """One directed comparison through the whole oracle pipeline.

    -> (result dict with the fields of b200ani_anim_result, kept rows (n,7), stats)."""
    rows, stats = anim_align(R, Q, maxmatch=maxmatch)
    keep = np.ones(len(rows), bool) if nofilter else filter_1to1(rows)
    kept = rows[keep][:, :7]
    out = np.zeros(5, dtype=np.int64)
    a = np.ascontiguousarray(kept, dtype=np.int64)
    lib().orc_parse_delta_reduce(_p(a, ctypes.c_int64), ctypes.c_int64(len(a)), _p(out, ctypes.c_int64))
    res = dict(ref_aln_len=int(out[0]), qry_aln_len=int(out[1]), sum_aligned=int(out[2]), sum_weighted=int(out[3]),
               sim_errs=int(out[4]), n_alignments=int(len(a)), status=0 if len(a) else 1)
    return res, kept, stats


def anim_pairs_mt(genomes, pairs, threads: int = 0, maxmatch: bool = False):
    """Whole oracle pipeline for many directed pairs on `threads` OpenMP threads (0 = all), no Python in the loop.
    genomes: list of OGenome; pairs: (m, 2) indices into it.  -> (array (m, 6): ref_aln_len, qry_aln_len, sum_aligned,
    sum_weighted, sim_errs, n_alignments; threads used)."""
    L = lib()
    L.orc_anim_pairs_mt.restype = ctypes.c_int
    L.orc_anim_pairs_mt.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    handles = (ctypes.c_void_p * len(genomes))(*[g._h for g in genomes])
    p = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
    out = np.zeros((len(p), 6), dtype=np.int64)
    used = L.orc_anim_pairs_mt(handles, p.ctypes.data, len(p), int(bool(maxmatch)), int(threads), out.ctypes.data)
    return out, int(used)


def anib_pairs_mt(genomes, pairs, threads: int = 0, fragsize: int = 1020):
    """ANIb counterpart of :func:`anim_pairs_mt`; pairs: (m, 2) (query, subject) indices.  -> (array (m, 6): aln_length, sim_errs,
    pid_milli_sum, n_frags, n_pass, n_hsps; threads used)."""
    L = lib()
    L.orc_anib_pairs_mt.restype = ctypes.c_int
    L.orc_anib_pairs_mt.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    handles = (ctypes.c_void_p * len(genomes))(*[g.anib_handle for g in genomes])
    p = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
    out = np.zeros((len(p), 6), dtype=np.int64)
    used = L.orc_anib_pairs_mt(handles, p.ctypes.data, len(p), int(fragsize), int(threads), out.ctypes.data)
    return out, int(used)


def anib_results_mt(genomes, pairs, threads: int = 0, fragsize: int = 1020):
    """:func:`anib_pair` result dicts for many (query, subject) index pairs into ``genomes``, computed on host threads
    (the word stage makes a 4 Mb pair take ~10 s on one core)."""
    out, _ = anib_pairs_mt(genomes, pairs, threads or min(len(pairs), os.cpu_count() or 1), fragsize)
    return [dict(aln_length=int(o[0]), sim_errs=int(o[1]), pid_milli_sum=int(o[2]), n_frags=int(o[3]), n_pass=int(o[4]), n_hsps=int(o[5]),
                 status=0 if o[4] else 1) for o in out]


# ----------------------------------------------------------------------------- ANIb reduction (anib.py:569-667)
def parse_blast_tab_reduce(rows: np.ndarray):
    """(aln_length, sim_errors, mean_pid) exactly as anib.parse_blast_tab(mode="ANIb").

    rows: one row per BLAST HSP in file order: (fragment number, blast_alnlen, blast_mismatch, blast_pid, qlen,
    blast_gaps) -- columns 1,3,4,5,7,15 of the reference's outfmt (anib.py:468-469).
    Restates anib.py:641-665: ani_alnlen = alnlen - gaps; ani_alnids = ani_alnlen - mismatch;
    keep rows with ani_alnlen/qlen > 0.7 and ani_alnids/qlen > 0.3 (strict); first kept row per fragment in
    file order; mean of their blast_pid (0 if none, anib.py:661-663), sums of ani_alnlen and of mismatch+gaps."""
    rows = np.asarray(rows, dtype=np.float64).reshape(-1, 6)
    frag, alnlen, mism, pid, qlen, gaps = (rows[:, i] for i in range(6))
    ani_alnlen = alnlen - gaps
    ani_alnids = ani_alnlen - mism
    keep = (ani_alnlen / qlen > 0.7) & (ani_alnids / qlen > 0.3)
    idx = np.flatnonzero(keep)
    _, first = np.unique(frag[idx], return_index=True)  # groupby(index).first(): groups sorted by fragment id
    sel = idx[first]
    if len(sel) == 0:
        return 0, 0, 0.0
    mean_pid = float(np.mean(pid[sel]))  # pandas' mean == numpy pairwise summation over the selected rows
    return int(ani_alnlen[sel].sum()), int((mism[sel] + gaps[sel]).sum()), mean_pid


# ----------------------------------------------------------------------------- ANIb aligner (oracle_anim.c)
def anib_pair(Q: OGenome, S: OGenome, fragsize: int = 1020, cap: int = 1 << 15):
    """Fragments of Q against S: -> (result dict, per-fragment rows (n,8), stats).

    result: aln_length, sim_errs, pid_milli_sum, n_frags, n_pass, n_hsps; mean_pid (percent) = pid_milli_sum / n_pass / 1000
    rows: frag, length, mismatch, gaps, nident, qlen, score, pid_milli  (one row per fragment that passes anib.py:646)"""
    L = lib()
    L.orc_anib_pair.restype = ctypes.c_int64
    L.orc_anib_pair.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]
    out = np.zeros(6, dtype=np.int64)
    rows = np.zeros((cap, 8), dtype=np.int64)
    stats = np.zeros(ANIM_NSTATS, dtype=np.int64)
    n = L.orc_anib_pair(Q.anib_handle, S.anib_handle, int(fragsize), out.ctypes.data, rows.ctypes.data, cap, stats.ctypes.data)
    assert 0 <= n <= cap
    res = dict(aln_length=int(out[0]), sim_errs=int(out[1]), pid_milli_sum=int(out[2]), n_frags=int(out[3]), n_pass=int(out[4]),
               n_hsps=int(out[5]), status=0 if out[4] else 1)
    return res, rows[:n].copy(), {k: v for k, v in zip(ANIM_STAT_NAMES, stats.tolist()) if k}
