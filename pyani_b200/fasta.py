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
    data = Path(path).read_bytes()
    ids: List[str] = []
    parts: List[bytes] = []
    lens: List[int] = []
    for block in data.split(b">")[1:]:
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
