"""ctypes binding of include/oov_b200.h — the only way the Python host reaches the GPU.

There is no CPU fallback: if the shared library is missing, cannot be loaded, or
the device is not sm_100, every op raises.  Error codes from the C-ABI are mapped
to ValueError (bad argument / alignment / workspace) or RuntimeError (CUDA, arch).
"""
from __future__ import annotations

import ctypes as C
import os

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG_DIR, "liboov_b200.so")

OOV_OK, OOV_ERR_ARG, OOV_ERR_ALIGN, OOV_ERR_ARCH, OOV_ERR_CUDA, OOV_ERR_WORKSPACE = 0, -1, -2, -3, -4, -5
OOV_F32, OOV_BF16 = 0, 1
PATH_AUTO, PATH_SIMT_FP32, PATH_TCGEN05 = 0, 1, 2
INT64_MAX = (1 << 63) - 1

c_i32, c_i64, c_u64, c_f32, c_vp, c_sz = C.c_int32, C.c_int64, C.c_uint64, C.c_float, C.c_void_p, C.c_size_t


class OovRows(C.Structure):
    """struct oov_rows (include/oov_b200.h)."""
    _fields_ = [
        ("ids", c_vp), ("ids_stride", c_i64), ("n", c_i64), ("n_old", c_i64), ("prime_pad", c_i64),
        ("iv_table", c_vp), ("iv_dtype", c_i32), ("out_dtype", c_i32), ("out", c_vp), ("out_stride", c_i64),
        ("D", c_i32), ("_pad", c_i32),
    ]


class OovDheNet(C.Structure):
    """struct oov_dhe_net (include/oov_b200.h)."""
    _fields_ = [("w", c_vp * 4), ("b", c_vp * 4), ("H", c_i32), ("hidden", c_i32), ("D", c_i32), ("F", c_i32)]


# name -> (restype, argtypes); must list EVERY symbol include/oov_b200.h declares
# (tests/test_abi.py cross-checks this table against the header).
SIGNATURES = {
    "oov_version": (C.c_char_p, []),
    "oov_last_error": (C.c_char_p, []),
    "oov_check_device": (c_i32, [c_i32]),
    "oov_launch_count": (c_u64, []),
    "oov_lsh_bits": (c_i32, [c_vp, c_i64, c_i32, c_vp, c_i32, c_vp, c_i64, c_i64, c_i64, c_f32, c_vp, c_vp, c_i32, c_vp]),
    "oov_lsh_embed": (c_i32, [c_vp, c_i64, c_i32, c_vp, c_i32, c_vp, c_i32, C.POINTER(OovRows), c_f32, c_vp, c_vp,
                              c_vp, c_sz, c_i32, c_vp]),
    "oov_lsh_embed_cast": (c_i32, [c_vp, c_i64, c_i32, c_vp, c_i32, c_vp, c_i32, C.POINTER(OovRows), c_f32, c_vp, c_vp,
                                   c_vp, c_sz, c_i32, c_vp, c_vp, c_i64, c_vp]),
    "oov_lsh_embed_workspace": (c_sz, [c_i64, c_i32, c_i32, c_i32]),
    "oov_slsh_embed": (c_i32, [c_vp, c_i64, c_i32, c_vp, c_i32, c_i32, c_vp, c_i32, C.POINTER(OovRows), c_f32, c_vp,
                               c_vp, c_vp]),
    "oov_dhe_hash": (c_i32, [c_vp, c_i64, c_i64, c_vp, c_i32, c_u64, c_vp, c_vp]),
    "oov_dhe_mlp": (c_i32, [c_vp, c_i64, C.POINTER(OovDheNet), c_vp, c_i32, c_i64, c_vp, c_sz, c_i32, c_vp]),
    "oov_dhe_embed": (c_i32, [c_vp, c_u64, C.POINTER(OovDheNet), C.POINTER(OovRows), c_vp, c_sz, c_i32, c_vp]),
    "oov_dhe_workspace": (c_sz, [c_i64, C.POINTER(OovDheNet), c_i32]),
    "oov_dhe_planes_ld": (c_i64, [c_i32]),
    "oov_dhe_hash_planes": (c_i32, [c_vp, c_i64, c_i64, c_vp, c_i32, c_u64, c_vp, c_vp]),
    "oov_dhe_embed_planes": (c_i32, [c_vp, C.POINTER(OovDheNet), C.POINTER(OovRows), c_vp, c_sz, c_vp]),
    "oov_fdhe_embed": (c_i32, [c_vp, c_u64, C.POINTER(OovDheNet), c_vp, c_i64, C.POINTER(OovRows), c_vp, c_sz, c_i32, c_vp]),
    "oov_fdhe_workspace": (c_sz, [c_i64, C.POINTER(OovDheNet), c_i32]),
    "oov_tc_linear": (c_i32, [c_vp, c_i64, c_vp, c_i64, c_i64, c_i32, c_i32, c_vp, c_i32, c_vp, c_i32, c_i64, c_vp]),
    "oov_col_mean": (c_i32, [c_vp, c_i32, c_i64, c_i32, c_vp, c_vp, c_sz, c_vp]),
    "oov_col_mean_workspace": (c_sz, [c_i64, c_i32]),
    "oov_const_embed": (c_i32, [c_vp, C.POINTER(OovRows), c_vp]),
    "oov_gather_rows": (c_i32, [c_vp, c_i32, c_i64, c_i32, c_vp, c_i64, c_i64, c_i64, c_vp, c_i32, c_i64, c_vp]),
    "oov_fullsort_topk": (c_i32, [c_vp, c_vp, c_i32, c_i64, c_i64, c_i32, c_i32, c_i64, c_i32, c_i64, c_i64, c_vp, c_vp,
                                  c_vp, c_vp, c_vp, c_sz, c_i32, c_vp]),
    "oov_fullsort_topk_workspace": (c_sz, [c_i64, c_i64, c_i32, c_i32, c_i32]),
    "oov_fullsort_scores": (c_i32, [c_vp, c_vp, c_i32, c_i64, c_i64, c_i32, c_i64, c_i32, c_i64, c_i64, c_vp, c_vp,
                                    c_vp, c_i64, c_vp]),
    "oov_dense_topk": (c_i32, [c_vp, c_i64, c_i64, c_i64, c_i32, c_vp, c_vp, c_vp]),
    "oov_topk_merge": (c_i32, [c_vp, c_vp, c_i32, c_i64, c_i32, c_vp, c_vp, c_vp]),
    "oov_fullsort_topk_keys": (c_i32, [c_vp, c_vp, c_i32, c_i64, c_i64, c_i32, c_i32, c_i32, c_i64, c_i64, c_vp, c_vp,
                                       c_i64, c_i64, c_i64, c_vp, c_vp, c_sz, c_i32, c_vp]),
    "oov_topk_merge_keys": (c_i32, [c_vp, c_i32, c_i64, c_i32, c_vp, c_vp, c_vp]),
    "oov_topk_hits": (c_i32, [c_vp, c_i64, c_i32, c_vp, c_vp, c_vp, c_vp]),
    "oov_topk_hits_collectors": (c_i32, [c_vp, c_vp, c_vp, c_i64, c_i32, c_vp, c_i64, c_i64, c_vp, c_vp, c_i32, c_vp, c_vp]),
    "oov_pairs_to_csr": (c_i32, [c_vp, c_vp, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp]),
    "oov_token_gather": (c_i32, [c_vp, c_i64, c_i32, c_vp, c_vp, c_i32, c_i64, c_i32, c_i64, c_i64, c_i32, c_i32,
                                 c_vp, c_vp, c_vp, c_i32, c_vp]),
    "oov_first_order_sum": (c_i32, [c_vp, c_i64, c_i32, c_vp, c_vp, c_i64, c_i64, c_i64, c_i32, c_i32, c_vp, c_vp,
                                    c_vp, c_vp]),
    "oov_map_ids": (c_i32, [c_vp, c_i64, c_i64, c_i64, c_i32, c_vp, c_vp]),
    "oov_cross_update": (c_i32, [c_vp, c_vp, c_vp, c_i64, c_vp, c_vp]),
    "oov_scatter_add_rows": (c_i32, [c_vp, c_i64, c_vp, c_i64, c_i64, c_i64, c_i64, c_i32, c_vp, c_vp]),
    "oov_lsh_embed_backward": (c_i32, [c_vp, c_i32, c_vp, c_i64, c_vp, c_i64, c_i64, c_i64, c_i32, c_vp, c_vp, c_sz, c_vp]),
    "oov_lsh_embed_backward_workspace": (c_sz, [c_i64]),
    "oov_fdhe_input": (c_i32, [c_vp, c_u64, c_i32, c_vp, c_i64, c_i32, c_vp, c_i64, c_i64, c_i64, c_vp, c_vp, c_sz, c_vp]),
    "oov_fdhe_input_workspace": (c_sz, [c_i64, c_i32]),
    "oov_linear_f32": (c_i32, [c_vp, c_vp, c_i64, c_i32, c_i32, c_vp, c_i32, c_vp, c_vp]),
    "oov_act": (c_i32, [c_vp, c_vp, c_i32, c_i64, c_i32, c_vp, c_i64, c_i64, c_vp, c_vp]),
    "oov_cin_outer": (c_i32, [c_vp, c_i64, c_i64, c_i64, c_i32, c_vp, c_i64, c_i64, c_i64, c_i32, c_i64, c_i32, c_vp, c_i64, c_vp]),
    "oov_cin_pool_dot": (c_i32, [c_vp, c_i64, c_i32, c_i32, c_i64, c_i32, c_vp, c_f32, c_i32, c_vp, c_vp]),
    "oov_cin_layer_supported": (c_i32, [c_i32, c_i32, c_i32, c_i32, c_i64]),
    "oov_cin_layer": (c_i32, [c_vp, c_i64, c_i32, c_vp, c_i64, c_i32, c_i32, c_i64, c_i32, c_vp, c_i64, c_vp, c_i32,
                              c_vp, c_i64, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp]),
    "oov_pair_topk": (c_i32, [c_vp, c_i32, c_vp, c_i32, c_i32, c_vp, c_vp, c_i64, c_i64, c_i32, c_i32, c_i64, c_i64, c_vp, c_i32, c_vp, c_vp, c_vp]),
}

_lib = None


def load() -> C.CDLL:
    """Load liboov_b200.so and bind every exported symbol.  Raises if the library is absent
    (run `python -c "import __graft_entry__ as g; g.build()"`); never falls back to the CPU."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: the CUDA extension has not been built and there is no CPU fallback. "
            "Run __graft_entry__.build() (needs nvcc).")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error() -> str:
    return load().oov_last_error().decode("utf-8", "replace")


def check(rc: int) -> None:
    if rc == OOV_OK:
        return
    msg = last_error()
    if rc in (OOV_ERR_ARG, OOV_ERR_ALIGN, OOV_ERR_WORKSPACE):
        raise ValueError(f"oov_b200 [{rc}]: {msg}")
    raise RuntimeError(f"oov_b200 [{rc}]: {msg}")
