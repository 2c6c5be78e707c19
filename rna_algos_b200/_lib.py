"""ctypes loader of the in-tree C-ABI library (include/rna_algos_b200.h).

The library is the product: if it is missing this module raises — there is no Python / CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

from .tables import AlignTables, ContraTables, TurnerTables

_HERE = os.path.dirname(os.path.abspath(__file__))
# RNA_B200_LIB: development override (A/B builds of the same library); the product is the in-tree library
LIB_PATH = os.environ.get("RNA_B200_LIB") or os.path.join(_HERE, "librna_algos_b200.so")

u8p = C.POINTER(C.c_uint8)
u16p = C.POINTER(C.c_uint16)
u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)
f32p = C.POINTER(C.c_float)
vp = C.c_void_p

RNA_OK = 0
STATUS = {
    0: "RNA_OK", 1: "RNA_ERR_BAD_ARG", 2: "RNA_ERR_INVALID_BASE", 3: "RNA_ERR_EMPTY_SEQ", 4: "RNA_ERR_TOO_LONG",
    5: "RNA_ERR_NO_TABLES", 6: "RNA_ERR_BAD_TABLES", 7: "RNA_ERR_CUDA", 8: "RNA_ERR_NO_DEVICE", 9: "RNA_ERR_NOMEM",
}
MODEL_TURNER, MODEL_CONTRA = 0, 1
NUMERIC_REF_EXACT, NUMERIC_FAST_F32, NUMERIC_FAST_F64 = 0, 1, 2


class FoldBatchDev(C.Structure):
    """RnaFoldBatchDev"""

    _fields_ = [
        ("h_offsets", vp), ("d_bases", vp), ("d_offsets", vp), ("d_bpp_offsets", vp),
        ("n_seqs", C.c_uint32), ("total_len", C.c_uint32), ("max_len", C.c_uint32),
        ("model", C.c_int), ("allows_short_hairpins", C.c_int),
        ("d_gammas", vp), ("n_gammas", C.c_uint32),
        ("d_out_logz", vp), ("d_out_bpp", vp), ("d_out_structs", vp), ("d_out_expect_acc", vp),
        ("d_out_pairs", vp), ("d_out_num_pairs", vp),
        ("d_out_sums", vp), ("d_sums_offsets", vp), ("inside_only", C.c_int),
    ]


class DurbinBatchDev(C.Structure):
    """RnaDurbinBatchDev"""

    _fields_ = [
        ("h_offsets", vp), ("h_pairs", vp), ("d_bases", vp), ("d_offsets", vp), ("d_pairs", vp),
        ("d_prob_offsets", vp), ("n_seqs", C.c_uint32), ("n_pairs", C.c_uint32), ("max_len", C.c_uint32),
        ("d_out_probs", vp),
    ]


class CallStats(C.Structure):
    _fields_ = [("kernel_launches", C.c_uint64), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64)]


_SIGS = {
    "rna_version": (C.c_char_p, []),
    "rna_sizeof_turner_tables": (C.c_size_t, []),
    "rna_sizeof_contra_tables": (C.c_size_t, []),
    "rna_sizeof_align_tables": (C.c_size_t, []),
    "rna_contra_tables_accumulate": (None, [C.POINTER(ContraTables)]),
    "rna_align_tables_contralign_v201": (None, [C.POINTER(AlignTables)]),
    "rna_create": (C.c_int, [C.c_int, C.POINTER(vp)]),
    "rna_destroy": (C.c_int, [vp]),
    "rna_last_error": (C.c_char_p, [vp]),
    "rna_device": (C.c_int, [vp]),
    "rna_set_numeric_mode": (C.c_int, [vp, C.c_int]),
    "rna_get_numeric_mode": (C.c_int, [vp]),
    "rna_validate_fold_lengths": (C.c_int, [vp, C.c_uint32]),
    "rna_set_turner_tables": (C.c_int, [vp, C.POINTER(TurnerTables)]),
    "rna_set_contra_tables": (C.c_int, [vp, C.POINTER(ContraTables)]),
    "rna_set_align_tables": (C.c_int, [vp, C.POINTER(AlignTables)]),
    "rna_mccaskill_centroid_batch": (C.c_int, [vp, vp, vp, C.c_uint32, C.c_int, C.c_int, vp, C.c_uint32, vp, vp,
                                               vp, vp, vp]),
    "rna_mccaskill_batch": (C.c_int, [vp, vp, vp, C.c_uint32, C.c_int, C.c_int, vp, vp, vp]),
    "rna_fold_sums_batch": (C.c_int, [vp, vp, vp, C.c_uint32, C.c_int, C.c_int, vp, vp, vp]),
    "rna_twoloop_scores": (C.c_int, [vp, vp, C.c_uint32, C.c_int, C.c_int, vp, C.c_uint64, vp]),
    "rna_centroid_batch": (C.c_int, [vp, vp, vp, vp, C.c_uint32, vp, C.c_uint32, vp, vp]),
    "rna_durbin_batch": (C.c_int, [vp, vp, vp, C.c_uint32, vp, C.c_uint32, vp, vp]),
    "rna_mccaskill_algo": (C.c_int, [vp, vp, C.c_uint32, C.c_int, C.c_int, vp, vp]),
    "rna_centroid_fold": (C.c_int, [vp, vp, C.c_uint32, C.c_float, vp, vp, vp, vp]),
    "rna_durbin_algo": (C.c_int, [vp, vp, C.c_uint32, vp, C.c_uint32, vp]),
    "rna_mccaskill_centroid_batch_dev": (C.c_int, [vp, C.POINTER(FoldBatchDev), vp]),
    "rna_durbin_batch_dev": (C.c_int, [vp, C.POINTER(DurbinBatchDev), vp]),
    "rna_validate_bases": (C.c_int, [vp, vp, C.c_uint32]),
    "rna_partition_lpt": (C.c_int, [vp, C.c_uint32, C.c_uint32, vp]),
    "rna_get_stats": (C.c_int, [vp, C.POINTER(CallStats)]),
    "rna_queue_create": (C.c_int, [vp, C.POINTER(vp)]),
    "rna_queue_destroy": (C.c_int, [vp]),
    "rna_queue_mccaskill_algo": (C.c_int, [vp, vp, C.c_uint32, C.c_int, C.c_int, vp, vp, C.c_float, vp, vp]),
    "rna_queue_stats": (C.c_int, [vp, u64p, u64p]),
    "rna_multi_create": (C.c_int, [vp, C.c_int, C.POINTER(vp)]),
    "rna_multi_destroy": (C.c_int, [vp]),
    "rna_multi_num_devices": (C.c_int, [vp]),
    "rna_multi_handle": (vp, [vp, C.c_int]),
    "rna_multi_last_error": (C.c_char_p, [vp]),
    "rna_multi_set_turner_tables": (C.c_int, [vp, C.POINTER(TurnerTables)]),
    "rna_multi_set_contra_tables": (C.c_int, [vp, C.POINTER(ContraTables)]),
    "rna_multi_set_align_tables": (C.c_int, [vp, C.POINTER(AlignTables)]),
    "rna_multi_set_numeric_mode": (C.c_int, [vp, C.c_int]),
    "rna_multi_mccaskill_centroid_batch": (C.c_int, [vp, vp, vp, C.c_uint32, C.c_int, C.c_int, vp, C.c_uint32, vp, vp,
                                                     vp, vp, vp]),
    "rna_multi_durbin_batch": (C.c_int, [vp, vp, vp, C.c_uint32, vp, C.c_uint32, vp, vp]),
    "rna_multi_last_shares": (C.c_int, [vp, vp, vp]),
}

_lib = None


def load() -> C.CDLL:
    """Load librna_algos_b200.so (built in-tree by `make` / __graft_entry__.build()).  Raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `make` (nvcc, sm_100a). "
            "rna_algos_b200 has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    assert lib.rna_sizeof_turner_tables() == C.sizeof(TurnerTables)
    assert lib.rna_sizeof_contra_tables() == C.sizeof(ContraTables)
    assert lib.rna_sizeof_align_tables() == C.sizeof(AlignTables)
    _lib = lib
    return lib


def exported_names():
    return list(_SIGS.keys())
