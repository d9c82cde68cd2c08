"""pyani_b200 -- B200-native engine for pyani's pairwise-comparison hot path (ANIm / ANIb / TETRA).

The product is the CUDA library ``libb200ani.so`` (sources in ``csrc/``, C ABI in ``include/b200ani.h``)
plus this thin host-side mirror of the reference's interfaces for the path:

* :mod:`pyani_b200.tetra`  mirrors ``pyani/tetra.py`` (calculate_tetra_zscore(s), calculate_correlations)
* :mod:`pyani_b200.anim`   mirrors the compute half of ``pyani/anim.py`` (parse_delta-shaped results,
  process_deltadir-shaped ANIResults tables) behind a ``run_dependency_graph``-style backend
* :mod:`pyani_b200.dist`   shards directed pairs over one-process-per-GPU ranks (torch.distributed)

There is no CPU fallback: importing works anywhere, but creating an :class:`Engine` without the built
library and a CUDA device raises.
"""
from .engine import Engine, EngineError, MultiEngine, lib_path  # noqa: F401

__version__ = "0.1.0"
