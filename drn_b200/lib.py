"""ctypes binding of libdrn_sm100.so (C ABI: include/drn_b200.h).

The library is built in-tree by `__graft_entry__.build()` / `make -C drn_b200/csrc`.  There is no fallback: if the
shared object is missing, or a call fails, a RuntimeError carrying `drn_last_error()` is raised.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdrn_sm100.so")

MAX_TAPS = 4
GEMM_ROWS, GEMM_WGRAD = 0, 2
OUT_STORE, OUT_ADD, OUT_ATOMIC = 0, 1, 2


class Planes(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("plane_stride", C.c_int64),
                ("B", C.c_int32), ("T", C.c_int32), ("P", C.c_int32), ("C", C.c_int32)]


class GemmDesc(C.Structure):
    _fields_ = [
        ("form", C.c_int32), ("b_mn", C.c_int32),
        ("a", Planes), ("b", Planes),
        ("B", C.c_int32), ("T", C.c_int32), ("N", C.c_int32), ("K", C.c_int32), ("M", C.c_int32),
        ("ntaps", C.c_int32),
        ("tap_shift", C.c_int32 * MAX_TAPS), ("tap_par", C.c_int32 * MAX_TAPS), ("tap_w", C.c_int32 * MAX_TAPS),
        ("a_c0", C.c_int32), ("b_c0", C.c_int32),
        ("nprod", C.c_int32), ("split_k", C.c_int32),
        ("out", C.c_void_p), ("out_ld", C.c_int64), ("out_col0", C.c_int32), ("out_mode", C.c_int32),
        ("out_tap_stride", C.c_int64), ("out_split_stride", C.c_int64),
        ("out_T", C.c_int32), ("out_t_mul", C.c_int32), ("out_t_add", C.c_int32),
        ("bias", C.c_void_p),
        ("rowscale", C.c_void_p), ("rowscale_ld", C.c_int32),
        ("out2", C.c_void_p), ("out2_ld", C.c_int64),
        ("outp", C.c_void_p), ("outp_ld", C.c_int64), ("outp_col0", C.c_int32), ("outp_plane_stride", C.c_int64),
        ("engine", C.c_int32),
        ("dbg_lbo", C.c_int32), ("dbg_sbo", C.c_int32), ("dbg_kadv", C.c_int32),
        ("stats", C.c_void_p),
    ]


class BnPart(C.Structure):
    _fields_ = [("c0", C.c_int32), ("n", C.c_int32), ("gamma", C.c_void_p), ("beta", C.c_void_p),
                ("running_mean", C.c_void_p), ("running_var", C.c_void_p), ("num_batches_tracked", C.c_void_p),
                ("dgamma", C.c_void_p), ("dbeta", C.c_void_p)]


class BnJob(C.Structure):
    """drn_bn_job_t: one BatchNorm application (conv block x pyramid level) of a multi-job launch."""
    _fields_ = [("y", C.c_void_p), ("y2", C.c_void_p), ("B", C.c_int32), ("T", C.c_int32), ("C", C.c_int32), ("nparts", C.c_int32),
                ("parts", BnPart * 2),
                ("coef", C.c_void_p), ("sums", C.c_void_p), ("counter", C.c_void_p), ("bcoef", C.c_void_p),
                ("up", C.c_void_p), ("up_plane_stride", C.c_int64), ("gate", C.c_void_p),
                ("out_a", C.c_void_p), ("a_plane_stride", C.c_int64), ("out_qa", C.c_void_p), ("qa_plane_stride", C.c_int64),
                ("da", C.c_void_p), ("dy", C.c_void_p), ("dy_plane_stride", C.c_int64),
                ("partials", C.c_void_p), ("partial_rows", C.c_int64)]


class HeadLevels(C.Structure):
    """drn_head_levels_t: per-level operands of the fused head projections."""
    _fields_ = [("nlevels", C.c_int32), ("B", C.c_int32), ("F", C.c_int32), ("T", C.c_int32 * 3),
                ("tower", C.c_void_p * 3), ("tower_plane_stride", C.c_int64 * 3),
                ("iou_hidden", C.c_void_p * 3), ("iou_hidden_plane_stride", C.c_int64 * 3),
                ("d_tower", C.c_void_p * 3)]


class PackItem(C.Structure):
    _fields_ = [("src", C.c_void_p), ("planes", C.c_void_p), ("grad", C.c_void_p),
                ("O", C.c_int32), ("C", C.c_int32), ("k", C.c_int32), ("Ototal", C.c_int32), ("o0", C.c_int32),
                ("plane_stride", C.c_int64), ("nslices", C.c_int32), ("slice_stride", C.c_int64)]


class SgemmJob(C.Structure):
    _fields_ = [("A", C.c_void_p), ("sam", C.c_int64), ("sak", C.c_int64), ("B", C.c_void_p), ("sbk", C.c_int64), ("sbn", C.c_int64),
                ("C", C.c_void_p), ("ldc", C.c_int64), ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32), ("bias", C.c_void_p),
                ("store", C.c_int32)]


class LinearJob(C.Structure):
    _fields_ = [("x", C.c_void_p), ("ldx", C.c_int64), ("W", C.c_void_p), ("ldw", C.c_int64), ("bias", C.c_void_p),
                ("out", C.c_void_p), ("ldo", C.c_int64), ("B", C.c_int32), ("N", C.c_int32), ("K", C.c_int32), ("relu", C.c_int32)]


class Qe(C.Structure):
    """drn_qe_t (include/drn_b200.h): query encoder parameters, gradients, outputs and workspace."""
    _fields_ = [
        ("B", C.c_int32), ("L", C.c_int32), ("H", C.c_int32), ("E", C.c_int32), ("tok_ld", C.c_int32),
        ("tokens", C.c_void_p), ("lengths", C.c_void_p),
        ("emb", C.c_void_p),
        ("w_ih", C.c_void_p * 2), ("w_hh", C.c_void_p * 2), ("b_ih", C.c_void_p * 2), ("b_hh", C.c_void_p * 2),
        ("w1", C.c_void_p), ("b1", C.c_void_p),
        ("w2", C.c_void_p * 3), ("b2", C.c_void_p * 3),
        ("wa", C.c_void_p), ("ba", C.c_void_p),
        ("cmd", C.c_void_p * 3),
        ("dcmd", C.c_void_p * 3),
        ("g_emb", C.c_void_p),
        ("g_w_ih", C.c_void_p * 2), ("g_w_hh", C.c_void_p * 2), ("g_b_ih", C.c_void_p * 2), ("g_b_hh", C.c_void_p * 2),
        ("g_w1", C.c_void_p), ("g_b1", C.c_void_p),
        ("g_w2", C.c_void_p * 3), ("g_b2", C.c_void_p * 3),
        ("g_wa", C.c_void_p), ("g_ba", C.c_void_p),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
    ]


_lib = None


def load():
    """Load the shared object once.  Raises if it has not been built (no CPU / torch fallback exists)."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError("libdrn_sm100.so is not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "or `make -C drn_b200/csrc` (expected at %s)" % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        lib.drn_last_error.restype = C.c_char_p
        lib.drn_version.restype = C.c_int
        lib.drn_qe_workspace_bytes.restype = C.c_size_t
        lib.drn_gemm_workspace_bytes.restype = C.c_size_t
        _lib = lib
    return _lib


def check(rc, what=""):
    if rc != 0:
        raise RuntimeError("libdrn_sm100 %s failed (%d): %s" % (what, rc, load().drn_last_error().decode()))


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
