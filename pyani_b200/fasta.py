"""FASTA ingest for the engine: residues of all records concatenated + record offsets.

Mirrors what the reference gets from ``SeqIO.parse(path, "fasta")`` (tetra.py:98, pyani_files.py:141):
record order is file order, every symbol counts towards the length (ambiguity codes included),
case is preserved here and folded on the device.
"""
from __future__ import annotations

from pathlib import Path
from typing import List, Tuple

import numpy as np

FASTA_SUFFIXES = (".fna", ".fa", ".fasta", ".fas", ".fsa_nt")  # pyani_files.py:64,80-85


def read_fasta(path) -> Tuple[np.ndarray, np.ndarray, List[str]]:
    """-> (uint8 residues, uint64 offsets[n_records+1], record ids)."""
    return _parse(Path(path).read_bytes(), path)


def _parse(data: bytes, path="<bytes>") -> Tuple[np.ndarray, np.ndarray, List[str]]:
    ids: List[str] = []
    parts: List[bytes] = []
    lens: List[int] = []
    # records start at a LINE-INITIAL '>' only (Bio.SeqIO's FASTA model): a '>' inside a description is header text
    if data.startswith(b">"):
        data = b"\n" + data
    for block in data.split(b"\n>")[1:]:
        nl = block.find(b"\n")
        header = block if nl < 0 else block[:nl]
        body = b"" if nl < 0 else block[nl + 1 :]
        toks = header.split()
        ids.append(toks[0].decode("ascii", "replace") if toks else "")
        seq = body.translate(None, b"\r\n \t")
        parts.append(seq)
        lens.append(len(seq))
    if not ids:
        raise ValueError(f"{path}: no FASTA records")
    arr = np.frombuffer(b"".join(parts), dtype=np.uint8)
    off = np.zeros(len(lens) + 1, dtype=np.uint64)
    off[1:] = np.cumsum(np.asarray(lens, dtype=np.uint64))
    return arr, off, ids


def get_fasta_files(dirname) -> List[Path]:
    """Same discovery rule as pyani_files.get_fasta_files (pyani_files.py:59-85)."""
    d = Path(dirname)
    return sorted(p for p in d.iterdir() if p.is_file() and p.suffix in FASTA_SUFFIXES)


def ingest_iter(paths, threads: int = 0):
    """Like :func:`ingest`, but a generator: records come out in input order as soon as they are ready while the pool keeps
    parsing the following files, so that a consumer (staging on the device) overlaps with the parsing."""
    import hashlib
    import os
    from concurrent.futures import ThreadPoolExecutor

    def one(p):
        p = Path(p)
        data = p.read_bytes()
        arr, off, ids = _parse(data, p)
        return {"path": p, "stem": p.stem, "md5": hashlib.md5(data).hexdigest(), "length": int(off[-1]),
                "residues": arr, "offsets": off, "ids": ids}

    paths = list(paths)
    n = threads or min(32, (os.cpu_count() or 1))
    if n <= 1 or len(paths) <= 1:
        for p in paths:
            yield one(p)
        return
    with ThreadPoolExecutor(n) as ex:
        yield from ex.map(one, paths)


def ingest(paths, threads: int = 0):
    """Parallel ingest of many FASTA files: parse + md5 + lengths in a thread pool.

    The reference parses every genome three times with Biopython (`pyani_orm.add_run_genomes`,
    `pyani_orm.py:566-570`; `pyani_files.get_sequence_lengths`, `pyani_files.py:139-142`) and hashes the file a
    fourth time (`download.create_hash`, `download.py:585-606`); here each file is read once.  File reading,
    `bytes.translate` and `hashlib.md5` all release the GIL, so plain threads scale.

    -> list of dicts in input order: path, stem (the legacy genome label, `pyani_files.py:141`), md5 (of the file bytes,
    what `<stem>.md5` sidecars hold), length (sum of record lengths, ambiguity symbols included, `pyani_tools.py:293-299`),
    residues / offsets / ids (what `Engine.add_genome` takes).
    """
    import hashlib
    import os
    from concurrent.futures import ThreadPoolExecutor

    def one(p):
        p = Path(p)
        data = p.read_bytes()
        arr, off, ids = _parse(data, p)
        return {"path": p, "stem": p.stem, "md5": hashlib.md5(data).hexdigest(), "length": int(off[-1]),
                "residues": arr, "offsets": off, "ids": ids}

    paths = list(paths)
    n = threads or min(32, (os.cpu_count() or 1))
    if n <= 1 or len(paths) <= 1:
        return [one(p) for p in paths]
    with ThreadPoolExecutor(n) as ex:
        return list(ex.map(one, paths))
