"""ctypes binding of include/b200ani.h (the drop-in C ABI). Fails loudly if the library is missing."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "libb200ani.so"
N_STAGES = 24


class AnimResult(C.Structure):
    """struct b200ani_anim_result (48 bytes), fields = anim.parse_delta's return kept exact."""

    _fields_ = [
        ("ref_aln_len", C.c_int64),
        ("qry_aln_len", C.c_int64),
        ("sum_aligned", C.c_int64),
        ("sum_weighted", C.c_int64),
        ("sim_errs", C.c_int64),
        ("n_alignments", C.c_int32),
        ("status", C.c_int32),
    ]


class AnibResult(C.Structure):
    """struct b200ani_anib_result (40 bytes), fields = anib.parse_blast_tab's return kept exact."""

    _fields_ = [
        ("aln_length", C.c_int64),
        ("sim_errs", C.c_int64),
        ("pid_milli_sum", C.c_int64),
        ("n_frags", C.c_int32),
        ("n_pass", C.c_int32),
        ("n_hsps", C.c_int32),
        ("status", C.c_int32),
    ]


# name -> (restype, argtypes); tests check every symbol of include/b200ani.h is listed and exported
SIGNATURES = {
    "b200ani_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int]),
    "b200ani_destroy": (None, [C.c_void_p]),
    "b200ani_version": (C.c_char_p, []),
    "b200ani_last_error": (C.c_char_p, [C.c_void_p]),
    "b200ani_add_genome": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(C.c_int32)]),
    "b200ani_clear": (C.c_int, [C.c_void_p]),
    "b200ani_staged_bytes": (C.c_int64, [C.c_void_p]),
    "b200ani_reset_indexes": (C.c_int, [C.c_void_p]),
    "b200ani_dpx_peak": (C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
    "b200ani_genome_count": (C.c_int, [C.c_void_p]),
    "b200ani_genome_length": (C.c_int64, [C.c_void_p, C.c_int32]),
    "b200ani_tetra_counts": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "b200ani_tetra_zscores": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "b200ani_tetra_corr": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "b200ani_anim": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p]),
    "b200ani_anim_alignments": (C.c_int64, [C.c_void_p, C.c_int32, C.c_int32, C.c_int, C.c_int, C.c_void_p, C.c_int64]),
    "b200ani_anim_rows": (C.c_int64, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "b200ani_write_delta": (C.c_int, [C.c_char_p, C.c_char_p, C.c_char_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32,
                                     C.c_void_p, C.c_int64]),
    "b200ani_anib_rows": (C.c_int64, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "b200ani_index_build": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "b200ani_index_info": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_uint32)]),
    "b200ani_index_export": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "b200ani_index_adopt": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_uint32]),
    "b200ani_sync": (C.c_int, [C.c_void_p]),
    "b200ani_sketch": (C.c_int, [C.c_void_p, C.c_int32, C.c_int, C.c_void_p, C.POINTER(C.c_int)]),
    "b200ani_anib": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]),
    "b200ani_anib_fragments": (C.c_int64, [C.c_void_p, C.c_int32, C.c_int32, C.c_int, C.c_void_p, C.c_int64]),
    "b200ani_create_multi": (C.c_int, [C.POINTER(C.c_void_p), C.c_void_p, C.c_int]),
    "b200ani_destroy_multi": (None, [C.c_void_p]),
    "b200ani_multi_last_error": (C.c_char_p, [C.c_void_p]),
    "b200ani_multi_device_count": (C.c_int, [C.c_void_p]),
    "b200ani_multi_device": (C.c_void_p, [C.c_void_p, C.c_int]),
    "b200ani_multi_add_genome": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(C.c_int32)]),
    "b200ani_multi_clear": (C.c_int, [C.c_void_p]),
    "b200ani_multi_anim": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p]),
    "b200ani_multi_anib": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]),
    "b200ani_last_profile": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "b200ani_stage_name": (C.c_char_p, [C.c_int]),
}

_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(
                f"{LIB_PATH} is not built. Run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(needs nvcc). pyani_b200 has no CPU fallback."
            )
        lib = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError = ABI drift, fail loudly
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib
