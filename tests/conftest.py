"""pytest config: `gpu` marker + shared fixtures. Non-GPU tests must pass in the build container."""
import json
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"
FIXTURES = ROOT / "oracle" / "_ref" / "fixtures"  # reference genomes copied by tests/golden/make_golden.py


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def tetra_gold():
    return json.load(open(GOLDEN / "tetra_reference.json"))


@pytest.fixture(scope="session")
def engine():
    import pyani_b200

    eng = pyani_b200.Engine(0)
    yield eng
    eng.close()


def fasta_text_to_arrays(txt: str):
    """Inline FASTA text -> (residues uint8, offsets uint64)."""
    seqs = []
    for block in txt.split(">")[1:]:
        lines = block.split("\n")
        seqs.append("".join(l.strip() for l in lines[1:]))
    off = np.zeros(len(seqs) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(s) for s in seqs])
    return np.frombuffer("".join(seqs).encode(), dtype=np.uint8).copy(), off


def fixture_genomes(sub="sequences"):
    d = FIXTURES / sub
    return sorted(d.glob("*.fna")) if d.exists() else []
