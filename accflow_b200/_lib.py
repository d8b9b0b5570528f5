"""ctypes binding of include/accflow_b200.h (the C-ABI drop-in boundary).

There is deliberately no fallback: if the shared library is missing or fails to load, every
product entry point raises.  Build it with ``python -m accflow_b200.build`` (or
``__graft_entry__.build()``).
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ACCFLOW_LIB") or os.path.join(HERE, "libaccflow_b200.so")   # env override: A/B builds
MAX_SRC = 4
ABI_VERSION = 5
TILES_AUTO, TILES_FORWARD, TILES_REVERSE = 0, 1, 2      # accflow_conv_desc.tile_order
ACT_NONE, ACT_RELU, ACT_SIGMOID, ACT_TANH = 0, 1, 2, 3
EPI_STORE, EPI_GRU_ZR, EPI_GRU_Q, EPI_STORE_POOL, EPI_ROWSTATS, EPI_STORE_T = 0, 1, 2, 3, 4, 5

fp = C.c_void_p  # device pointers travel as integers


class ConvDesc(C.Structure):
    _fields_ = [
        ("src", fp * MAX_SRC), ("src_c", C.c_int * MAX_SRC), ("src_ld", C.c_int * MAX_SRC), ("nsrc", C.c_int),
        ("batch", C.c_int), ("in_h", C.c_int), ("in_w", C.c_int),
        ("weight", fp), ("weight_batch_stride", C.c_longlong),
        ("kh", C.c_int), ("kw", C.c_int), ("stride", C.c_int), ("pad_h", C.c_int), ("pad_w", C.c_int),
        ("cout", C.c_int), ("cout_pad", C.c_int),
        ("alpha", C.c_float), ("scale", fp), ("shift", fp),
        ("act", C.c_int), ("act_split", C.c_int), ("act2", C.c_int),
        ("residual", fp), ("res_ld", C.c_int), ("post_relu", C.c_int), ("epilogue", C.c_int),
        ("out", fp), ("out_ld", C.c_int), ("out2", fp), ("out2_ld", C.c_int),
        ("h", fp), ("h_ld", C.c_int), ("z", fp), ("z_ld", C.c_int), ("pool_w", C.c_int),
        ("pre_add", fp), ("pre_ld", C.c_int), ("pre_mod", C.c_int), ("row_stats", fp), ("out_h", C.c_int), ("out_w", C.c_int),
        ("tile_order", C.c_int),
    ]


class TcWeights(C.Structure):
    _fields_ = [("planes", fp), ("nplanes", C.c_int), ("rows", C.c_int), ("k", C.c_int), ("k_pitch", C.c_int),
                ("t", C.c_int), ("plane_stride", C.c_longlong)]


class TcIO(C.Structure):
    _fields_ = [("src_planes", fp * MAX_SRC), ("src_pitch", C.c_int * MAX_SRC), ("src_plane_stride", C.c_longlong * MAX_SRC),
                ("out_planes", fp), ("out_pitch", C.c_int), ("out_plane_stride", C.c_longlong),
                ("out2_planes", fp), ("out2_pitch", C.c_int), ("out2_plane_stride", C.c_longlong),
                ("h_planes", fp), ("h_pitch", C.c_int), ("h_plane_stride", C.c_longlong)]


i, ll, f = C.c_int, C.c_longlong, C.c_float
# name -> argtypes (restype is int unless listed in _RESTYPE); mirrors include/accflow_b200.h
SIGNATURES = {
    "accflow_abi_version": [],
    "accflow_sizeof": [i],
    "accflow_last_error": [C.c_char_p, C.c_size_t],
    "accflow_launch_count": [i],
    "accflow_launch_count_add": [ll],
    "accflow_conv2d_f32": [C.POINTER(ConvDesc), fp],
    "accflow_conv2d_tc": [C.POINTER(ConvDesc), C.POINTER(TcIO), C.POINTER(TcWeights), i, fp],
    "accflow_tc_debug_trace": [fp, i],
    "accflow_tc_rowstat_parts": [i, i],
    "accflow_softmax_stats_finalize": [fp, ll, i, fp, fp],
    "accflow_split_bf16_planes": [fp, ll, i, i, i, i, ll, i, fp, fp],
    "accflow_conv_smallc_f32": [fp, i, i, i, i, i, fp, fp, fp, i, i, i, i, fp, i, fp, i, ll, i, fp],
    "accflow_instnorm_chunks": [i],
    "accflow_instnorm_f32": [fp, i, i, i, f, i, fp, i, fp, fp, fp, fp],
    "accflow_instnorm_planes_f32": [fp, i, i, i, f, i, fp, i, fp, fp, fp, fp, i, ll, i, fp],
    "accflow_nhwc_transpose_f32": [fp, i, i, i, i, fp, i, fp],
    "accflow_corr_pool_f32": [fp, ll, i, i, fp, fp, fp, fp],
    "accflow_corr_lookup_f32": [fp, fp, fp, fp, i, i, i, i, fp, fp, i, fp, fp, i, fp, i, ll, fp, i, ll, i, fp],
    "accflow_stem_patch_planes": [fp, i, i, i, fp, i, ll, i, fp],
    "accflow_stem_rows_planes": [fp, i, i, i, fp, i, ll, i, fp],
    "accflow_flow_patch_f32": [fp, i, i, i, fp, i, fp, i, ll, i, fp],
    "accflow_conv3x3_smallcout_f32": [fp, i, i, i, i, i, fp, fp, fp, i, i, fp, i, fp],
    "accflow_coords_init_f32": [fp, i, i, i, fp, fp],
    "accflow_axpy_f32": [fp, fp, f, ll, fp],
    "accflow_tapsum3x3_f32": [fp, i, i, i, i, i, fp, fp, i, fp, i, fp, i, fp],
    "accflow_convex_upsample_f32": [fp, i, i, fp, i, i, i, i, fp, fp],
    "accflow_downflow8_f32": [fp, i, i, i, fp, fp],
    "accflow_upflow8_f32": [fp, i, i, i, i, fp, fp],
    "accflow_warp_occ_f32": [fp, i, fp, i, fp, i, i, i, i, fp, i, fp, i, fp],
    "accflow_backwarp_nchw_f32": [fp, fp, i, i, i, i, fp, fp],
    "accflow_epe_metrics_f32": [fp, fp, fp, i, i, i, fp, fp, fp],
    "accflow_deform_gather_f32": [fp, i, fp, i, i, i, i, i, fp, fp],
    "accflow_blend_f32": [fp, fp, fp, i, ll, i, fp, fp],
    "accflow_softmax_rows_f32": [fp, ll, i, fp],
}
_RESTYPE = {"accflow_launch_count": ll, "accflow_launch_count_add": ll}
_NO_CHECK = {"accflow_abi_version", "accflow_sizeof", "accflow_last_error", "accflow_launch_count", "accflow_launch_count_add",
             "accflow_instnorm_chunks", "accflow_tc_rowstat_parts"}

_lib = None


class AccflowError(RuntimeError):
    pass


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise AccflowError(f"{LIB_PATH} is missing: the CUDA extension is not built "
                           "(run `python -m accflow_b200.build`); there is no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.argtypes = argtypes
        fn.restype = _RESTYPE.get(name, C.c_int)
    if lib.accflow_abi_version() != ABI_VERSION:
        raise AccflowError("ABI version mismatch between _lib.py and libaccflow_b200.so")
    for which, mirror in enumerate((ConvDesc, TcWeights, TcIO)):
        if lib.accflow_sizeof(which) != C.sizeof(mirror):
            raise AccflowError(f"{mirror.__name__}: ctypes mirror is {C.sizeof(mirror)} bytes, the library's struct "
                               f"{lib.accflow_sizeof(which)} (include/accflow_b200.h and _lib.py disagree)")
    _lib = lib
    return lib


def last_error() -> str:
    buf = C.create_string_buffer(512)
    load().accflow_last_error(buf, 512)
    return buf.value.decode(errors="replace")


def call(name: str, *args):
    """Invoke a C-ABI entry point; non-zero return codes raise with the library's message."""
    rc = getattr(load(), name)(*args)
    if name not in _NO_CHECK and rc != 0:
        raise AccflowError(f"{name} failed (rc={rc}): {last_error()}")
    return rc
