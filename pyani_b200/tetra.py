"""TETRA on the B200 engine behind the reference's function names (pyani/tetra.py).

``calculate_tetra_zscore`` / ``calculate_tetra_zscores`` / ``calculate_correlations`` take and return
the same Python objects as the reference (dict of tetramer -> Z, DataFrame of correlations) and raise
the same ``AssertionError`` when two genomes observe different tetramer sets (tetra.py:174-175).
"""
from __future__ import annotations

from pathlib import Path
from typing import Dict, Iterable, Optional

import numpy as np

from .engine import Engine

TETRAMERS = ["".join("ACGT"[(t >> s) & 3] for s in (6, 4, 2, 0)) for t in range(256)]
_default_engine: Optional[Engine] = None


def _engine(engine: Optional[Engine]) -> Engine:
    global _default_engine
    if engine is not None:
        return engine
    if _default_engine is None:
        _default_engine = Engine(0)
    return _default_engine


def _to_dict(z_row: np.ndarray, obs_row: np.ndarray) -> Dict[str, float]:
    # the reference has no key for an unobserved tetramer (tetra.py:120)
    return {TETRAMERS[t]: float(z_row[t]) for t in range(256) if obs_row[t]}


def calculate_tetra_zscore(filename: Path, engine: Optional[Engine] = None) -> Dict[str, float]:
    """tetra.py:78 -- Z-scores of one FASTA file, as {tetramer: z}."""
    eng = _engine(engine)
    gid = eng.add_fasta(filename)
    try:
        z, obs = eng.tetra_zscores([gid])
    finally:
        if engine is None:  # the shared default engine never hands genome ids out: drop what this call staged
            eng.clear()
    return _to_dict(z[0], obs[0])


def calculate_tetra_zscores(infilenames: Iterable, engine: Optional[Engine] = None) -> Dict[str, Dict[str, float]]:
    """tetra.py:66 -- {file stem: {tetramer: z}}; one batched device call for all files."""
    eng = _engine(engine)
    files = [Path(f) for f in infilenames]
    try:
        gids = [eng.add_fasta(f) for f in files]
        z, obs = eng.tetra_zscores(gids)
    finally:
        if engine is None:
            eng.clear()
    return {f.stem: _to_dict(z[i], obs[i]) for i, f in enumerate(files)}


def calculate_correlations(tetra_z: Dict[str, Dict[str, float]], engine: Optional[Engine] = None):
    """tetra.py:158 -- DataFrame of Pearson correlations, organisms sorted, diagonal 1.0."""
    import pandas as pd

    eng = _engine(engine)
    orgs = sorted(tetra_z.keys())
    keysets = [sorted(tetra_z[o].keys()) for o in orgs]
    for ks in keysets[1:]:
        if ks != keysets[0]:
            raise AssertionError()  # tetra.py:174-175
    mask = np.zeros(256, dtype=np.uint8)
    z = np.zeros((len(orgs), 256), dtype=np.float64)
    index = {t: i for i, t in enumerate(TETRAMERS)}
    for k in keysets[0]:
        mask[index[k]] = 1
    for i, o in enumerate(orgs):
        for k, v in tetra_z[o].items():
            z[i, index[k]] = v
    corr = eng.tetra_corr(z, mask) if len(orgs) > 1 else np.ones((len(orgs), len(orgs)))
    return pd.DataFrame(corr, index=orgs, columns=orgs, dtype=float)
