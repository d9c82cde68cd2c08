"""Host-side handle on one GPU's b200ani context (one Engine per process / per GPU)."""
from __future__ import annotations

import ctypes as C
from pathlib import Path
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from .fasta import read_fasta


class EngineError(RuntimeError):
    pass


def lib_path() -> Path:
    return _lib.LIB_PATH


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class Engine:
    """Owns a ``b200ani_ctx``: genomes staged in HBM, TETRA and ANIm entry points.

    >>> eng = Engine(device=0)
    >>> gid = eng.add_fasta("genome.fna")
    >>> z, observed = eng.tetra_zscores([gid])
    """

    def __init__(self, device: int = 0):
        self._lib = _lib.load()
        h = C.c_void_p()
        rc = self._lib.b200ani_create(C.byref(h), int(device))
        if rc != 0 or not h:
            raise EngineError(f"b200ani_create(device={device}) failed with rc={rc}: a CUDA device is required (no CPU fallback)")
        self._h = h
        self.device = int(device)
        self.names: List[str] = []
        self.contig_ids: List[List[str]] = []
        self.contig_lens: List[np.ndarray] = []

    # -- lifetime
    def close(self):
        if getattr(self, "_h", None):
            self._lib.b200ani_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc: int, what: str):
        if rc != 0:
            raise EngineError(f"{what}: rc={rc}: {self._lib.b200ani_last_error(self._h).decode()}")

    @property
    def version(self) -> str:
        return self._lib.b200ani_version().decode()

    # -- genomes
    def add_genome(self, residues: np.ndarray, contig_off: np.ndarray, name: str = "", contig_ids: Optional[List[str]] = None) -> int:
        residues = np.ascontiguousarray(residues, dtype=np.uint8)
        contig_off = np.ascontiguousarray(contig_off, dtype=np.uint64)
        gid = C.c_int32(-1)
        rc = self._lib.b200ani_add_genome(self._h, _ptr(residues), _ptr(contig_off), len(contig_off) - 1, C.byref(gid))
        self._check(rc, "add_genome")
        self.names.append(name or f"genome{gid.value}")
        self.contig_ids.append(contig_ids or [str(i) for i in range(len(contig_off) - 1)])
        self.contig_lens.append(np.diff(contig_off.astype(np.int64)))
        return gid.value

    def add_fasta(self, path) -> int:
        arr, off, ids = read_fasta(path)
        return self.add_genome(arr, off, Path(path).stem, ids)

    def add_fasta_many(self, paths, threads: int = 0) -> List[int]:
        """Ingest many FASTA files with a thread pool (parse + md5 in parallel), stage them in input order.
        The per-file records (md5, length, ...) are kept in ``self.ingested`` for the caller's bookkeeping
        (pyani indexes genomes by md5, `pyani_orm.py:194,572`)."""
        from .fasta import ingest_iter

        gids, self.ingested = [], []
        # files are parsed (and hashed) by a thread pool and handed over in input order: while genome i goes through the
        # pinned double buffer to the device and is packed there (ctypes releases the GIL), the pool is already parsing i+1 ...
        for r in ingest_iter(paths, threads):
            gids.append(self.add_genome(r["residues"], r["offsets"], r["stem"], r["ids"]))
            self.ingested.append({k: v for k, v in r.items() if k not in ("residues", "offsets")})
        return gids

    def staged_bytes(self) -> int:
        """Residues that went through the context's pinned double buffer since the last clear()."""
        return int(self._lib.b200ani_staged_bytes(self._h))

    def clear(self):
        """Drop all staged genomes (ids restart at 0)."""
        self._check(self._lib.b200ani_clear(self._h), "clear")
        self.names, self.contig_ids, self.contig_lens = [], [], []
        self.__dict__.pop("_rec_tables", None)

    def reset_indexes(self):
        self._check(self._lib.b200ani_reset_indexes(self._h), "reset_indexes")

    def dpx_peak_gops(self) -> float:
        v = C.c_double(0)
        self._check(self._lib.b200ani_dpx_peak(self._h, C.byref(v)), "dpx_peak")
        return v.value

    def genome_length(self, gid: int) -> int:
        return int(self._lib.b200ani_genome_length(self._h, gid))

    def __len__(self):
        return int(self._lib.b200ani_genome_count(self._h))

    # -- TETRA
    def tetra_counts(self, gid: int):
        c2, c3, c4 = (np.zeros(k, dtype=np.int64) for k in (16, 64, 256))
        self._check(self._lib.b200ani_tetra_counts(self._h, gid, _ptr(c2), _ptr(c3), _ptr(c4)), "tetra_counts")
        return c2, c3, c4

    def tetra_zscores(self, gids: Sequence[int]):
        ids = np.ascontiguousarray(gids, dtype=np.int32)
        z = np.zeros((len(ids), 256), dtype=np.float64)
        obs = np.zeros((len(ids), 256), dtype=np.uint8)
        self._check(self._lib.b200ani_tetra_zscores(self._h, _ptr(ids), len(ids), _ptr(z), _ptr(obs)), "tetra_zscores")
        return z, obs

    def tetra_corr(self, z: np.ndarray, mask: Optional[np.ndarray] = None) -> np.ndarray:
        z = np.ascontiguousarray(z, dtype=np.float64)
        n = z.shape[0]
        mask_arr = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        corr = np.zeros((n, n), dtype=np.float64)
        self._check(self._lib.b200ani_tetra_corr(self._h, _ptr(z), None if mask_arr is None else _ptr(mask_arr), n, _ptr(corr)), "tetra_corr")
        return corr

    # -- ANIm
    def anim(self, pairs: np.ndarray, maxmatch: bool = False, nofilter: bool = False) -> np.ndarray:
        """pairs: (m,2) int32 of (ref genome id, qry genome id) -> structured array of AnimResult."""
        pairs = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
        out = np.zeros(len(pairs), dtype=ANIM_DTYPE)
        if len(pairs):
            self._check(self._lib.b200ani_anim(self._h, _ptr(pairs), len(pairs), int(maxmatch), int(nofilter), _ptr(out)), "anim")
        return out

    def anim_alignments(self, ref: int, qry: int, maxmatch: bool = False, nofilter: bool = False, cap: int = 1 << 16) -> np.ndarray:
        while True:
            rows = np.zeros((cap, 7), dtype=np.int64)
            n = self._lib.b200ani_anim_alignments(self._h, ref, qry, int(maxmatch), int(nofilter), _ptr(rows), cap)
            if n < 0:
                self._check(int(n), "anim_alignments")
            if n <= cap:  # never hand back a truncated list: a short .delta would parse to wrong lengths silently
                return rows[:n].copy()
            cap = int(n)

    def anim_rows(self, pairs: np.ndarray, maxmatch: bool = False, nofilter: bool = False, all_rows: bool = False):
        """All pairs in one device pass -> (results, rows (n,7) int64, row_start (m+1,)): rows[row_start[i]:row_start[i+1]] are the
        alignments of pair i (what its .filter -- with nofilter its .delta -- would hold), see :meth:`anim_alignments`.
        all_rows=True: every alignment is returned plus a fourth value, kept (n,) bool: the .delta rows and which of them are in
        the .filter, from the same pass."""
        pairs = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
        m = len(pairs)
        out = np.zeros(m, dtype=ANIM_DTYPE)
        start = np.zeros(m + 1, dtype=np.int64)
        cap = max(1024, 64 * m)
        while True:
            rows = np.zeros((cap, 7), dtype=np.int64)
            kept = np.zeros(cap, dtype=np.uint8) if all_rows else None
            n = self._lib.b200ani_anim_rows(self._h, _ptr(pairs), m, int(maxmatch), int(nofilter), _ptr(out), _ptr(rows), cap, _ptr(start),
                                            _ptr(kept) if all_rows else None)
            if n < 0:
                self._check(int(n), "anim_rows")
            if n <= cap:
                return (out, rows[:n].copy(), start, kept[:n].astype(bool)) if all_rows else (out, rows[:n].copy(), start)
            cap = int(n)

    def write_delta(self, path, ref_path, qry_path, ref_gid: int, qry_gid: int, rows: np.ndarray) -> None:
        """Native .delta / .filter emitter (b200ani_write_delta) for alignment rows of the pair (ref_gid, qry_gid)."""
        import ctypes as C

        rows = np.ascontiguousarray(rows, dtype=np.int64).reshape(-1, 7)
        ri, rl = self._record_tables(ref_gid)
        qi, ql = self._record_tables(qry_gid)
        Path(path).parent.mkdir(parents=True, exist_ok=True)
        rc = self._lib.b200ani_write_delta(str(path).encode(), str(ref_path).encode(), str(qry_path).encode(), ri, _ptr(rl), len(rl), qi, _ptr(ql), len(ql),
                                           _ptr(rows), len(rows))
        if rc != 0:
            raise EngineError(f"write_delta({path}): rc={rc}")

    def _record_tables(self, gid: int):
        """(C array of record ids, int64 record lengths) of a staged genome, built once."""
        import ctypes as C

        cache = self.__dict__.setdefault("_rec_tables", {})
        if gid not in cache:
            ids = self.contig_ids[gid]
            arr = (C.c_char_p * len(ids))(*[i.encode() for i in ids])
            cache[gid] = (arr, np.ascontiguousarray(self.contig_lens[gid], dtype=np.int64))
        return cache[gid]

    def anib_rows(self, pairs: np.ndarray, fragsize: int = 1020):
        """All (query, subject) pairs in one device pass -> (results, rows (n,8) int64, row_start (m+1,)), see :meth:`anib_fragments`."""
        pairs = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
        m = len(pairs)
        out = np.zeros(m, dtype=ANIB_DTYPE)
        start = np.zeros(m + 1, dtype=np.int64)
        cap = max(4096, 2048 * m)
        while True:
            rows = np.zeros((cap, 8), dtype=np.int64)
            n = self._lib.b200ani_anib_rows(self._h, _ptr(pairs), m, int(fragsize), _ptr(out), _ptr(rows), cap, _ptr(start))
            if n < 0:
                self._check(int(n), "anib_rows")
            if n <= cap:
                return out, rows[:n].copy(), start
            cap = int(n)

    # -- seed indexes shared between GPUs (pyani_b200.dist.share_indexes)
    def index_build(self, gids: Sequence[int]) -> None:
        ids = np.ascontiguousarray(gids, dtype=np.int32)
        if len(ids):
            self._check(self._lib.b200ani_index_build(self._h, _ptr(ids), len(ids)), "index_build")

    def index_info(self, gid: int) -> Tuple[int, int]:
        """(n_mini, table_bits) of a genome's valid seed index."""
        n, b = C.c_int64(0), C.c_uint32(0)
        self._check(self._lib.b200ani_index_info(self._h, int(gid), C.byref(n), C.byref(b)), "index_info")
        return n.value, b.value

    def index_export(self, gid: int, d_kmer: int, d_posflag: int, d_table: int, d_bitmap: int) -> None:
        """Copy the index arrays to device addresses of the caller (asynchronous: :meth:`sync` before anyone else reads them)."""
        self._check(self._lib.b200ani_index_export(self._h, int(gid), C.c_void_p(d_kmer), C.c_void_p(d_posflag), C.c_void_p(d_table), C.c_void_p(d_bitmap)),
                    "index_export")

    def index_adopt(self, gid: int, d_kmer: int, d_posflag: int, d_table: int, d_bitmap: int, n_mini: int, table_bits: int) -> None:
        """Use index arrays in caller-owned device memory (built on another GPU) for this genome; they must outlive their use."""
        self._check(self._lib.b200ani_index_adopt(self._h, int(gid), C.c_void_p(d_kmer), C.c_void_p(d_posflag), C.c_void_p(d_table), C.c_void_p(d_bitmap),
                                                  int(n_mini), int(table_bits)), "index_adopt")

    def sync(self) -> None:
        self._check(self._lib.b200ani_sync(self._h), "sync")

    def sketch(self, gid: int, s: int = 2048) -> np.ndarray:
        """Bottom-s MinHash sketch of a staged genome (ascending distinct uint32 hashes of its seed k-mers), see pyani_b200.prefilter."""
        out = np.zeros(s, dtype=np.uint32)
        n = C.c_int(0)
        self._check(self._lib.b200ani_sketch(self._h, int(gid), int(s), _ptr(out), C.byref(n)), "sketch")
        return out[: n.value].copy()

    # -- ANIb
    def anib(self, pairs: np.ndarray, fragsize: int = 1020) -> np.ndarray:
        """pairs: (m,2) int32 of (query genome id, subject genome id) -> structured array of AnibResult."""
        pairs = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
        out = np.zeros(len(pairs), dtype=ANIB_DTYPE)
        if len(pairs):
            self._check(self._lib.b200ani_anib(self._h, _ptr(pairs), len(pairs), int(fragsize), _ptr(out)), "anib")
        return out

    def anib_fragments(self, qry: int, sbj: int, fragsize: int = 1020, cap: int = 1 << 15) -> np.ndarray:
        """Kept HSP per passing fragment: rows (frag, length, mismatch, gaps, nident, qlen, score, pid_milli)."""
        while True:
            rows = np.zeros((cap, 8), dtype=np.int64)
            n = self._lib.b200ani_anib_fragments(self._h, qry, sbj, int(fragsize), _ptr(rows), cap)
            if n < 0:
                self._check(int(n), "anib_fragments")
            if n <= cap:
                return rows[:n].copy()
            cap = int(n)

    def last_profile(self) -> Tuple[Dict[str, float], Dict[str, int]]:
        ms = np.zeros(_lib.N_STAGES, dtype=np.float64)
        cnt = np.zeros(_lib.N_STAGES, dtype=np.int64)
        self._lib.b200ani_last_profile(self._h, _ptr(ms), _ptr(cnt))
        names = [self._lib.b200ani_stage_name(i).decode() for i in range(_lib.N_STAGES)]
        cnames = [self._lib.b200ani_stage_name(_lib.N_STAGES + i).decode() for i in range(_lib.N_STAGES)]
        return ({n: float(v) for n, v in zip(names, ms) if n}, {n: int(v) for n, v in zip(cnames, cnt) if n})


ANIM_DTYPE = np.dtype(
    [("ref_aln_len", "<i8"), ("qry_aln_len", "<i8"), ("sum_aligned", "<i8"), ("sum_weighted", "<i8"),
     ("sim_errs", "<i8"), ("n_alignments", "<i4"), ("status", "<i4")]
)
assert ANIM_DTYPE.itemsize == 48
ANIB_DTYPE = np.dtype(
    [("aln_length", "<i8"), ("sim_errs", "<i8"), ("pid_milli_sum", "<i8"), ("n_frags", "<i4"), ("n_pass", "<i4"),
     ("n_hsps", "<i4"), ("status", "<i4")]
)
assert ANIB_DTYPE.itemsize == 40


class MultiEngine:
    """Several GPUs of one node behind one handle (``b200ani_create_multi``): genomes replicated on every device, the pair list
    dealt over the devices, ONE NCCL all-gather of the result structs inside the library.  The alternative to one process
    per GPU (:mod:`pyani_b200.dist`) for callers that want a single process."""

    def __init__(self, devices: Sequence[int]):
        self._lib = _lib.load()
        ids = np.ascontiguousarray(devices, dtype=np.int32)
        h = C.c_void_p()
        rc = self._lib.b200ani_create_multi(C.byref(h), _ptr(ids), len(ids))
        if rc != 0 or not h:
            raise EngineError(f"b200ani_create_multi(devices={list(devices)}) failed with rc={rc} (CUDA devices and NCCL are required)")
        self._h = h
        self.devices = [int(d) for d in devices]
        self.names: List[str] = []

    def close(self):
        if getattr(self, "_h", None):
            self._lib.b200ani_destroy_multi(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, what: str):
        if rc != 0:
            raise EngineError(f"{what}: rc={rc}: {self._lib.b200ani_multi_last_error(self._h).decode()}")

    def add_genome(self, residues: np.ndarray, contig_off: np.ndarray, name: str = "") -> int:
        residues = np.ascontiguousarray(residues, dtype=np.uint8)
        contig_off = np.ascontiguousarray(contig_off, dtype=np.uint64)
        gid = C.c_int32(-1)
        self._check(self._lib.b200ani_multi_add_genome(self._h, _ptr(residues), _ptr(contig_off), len(contig_off) - 1, C.byref(gid)), "multi_add_genome")
        self.names.append(name or f"genome{gid.value}")
        return gid.value

    def clear(self):
        self._check(self._lib.b200ani_multi_clear(self._h), "multi_clear")
        self.names = []

    def anim(self, pairs: np.ndarray, maxmatch: bool = False, nofilter: bool = False) -> np.ndarray:
        pairs = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
        out = np.zeros(len(pairs), dtype=ANIM_DTYPE)
        if len(pairs):
            self._check(self._lib.b200ani_multi_anim(self._h, _ptr(pairs), len(pairs), int(maxmatch), int(nofilter), _ptr(out)), "multi_anim")
        return out

    def anib(self, pairs: np.ndarray, fragsize: int = 1020) -> np.ndarray:
        pairs = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
        out = np.zeros(len(pairs), dtype=ANIB_DTYPE)
        if len(pairs):
            self._check(self._lib.b200ani_multi_anib(self._h, _ptr(pairs), len(pairs), int(fragsize), _ptr(out)), "multi_anib")
        return out
