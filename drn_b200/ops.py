"""Python launchers for the kernels of libdrn_sm100.so.  Thin: argument marshalling only, no math."""
import ctypes as C

import torch

from . import lib as L
from .planes import Planes

import os

# numeric mode of the tensor-core contraction: 3 = split-BF16 parity mode (default: fp32-equivalent, every reported number),
# 1 = single BF16 pass (OPT-IN fast mode, DRN_NPROD=1: ~3x fewer tensor-core FLOPs, BF16-level error -- bench.py's
# extra.fast_mode reports its speed and its error against the oracle), 4 = + lo*lo
NPROD = int(os.environ.get("DRN_NPROD", "3"))
if NPROD not in (1, 3, 4):
    raise ValueError("DRN_NPROD must be 1, 3 or 4")
ENGINE = 0  # 0 = tcgen05 (product), 1 = fp32 CUDA-core checker (tests only)


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def desc(form, a, b, B, T, N, K=0, M=0, taps=((0, 0, 0),), b_mn=0, a_c0=0, b_c0=0, nprod=None, split_k=1,
         out=None, out_ld=None, out_col0=0, out_mode=L.OUT_STORE, out_tap_stride=0, out_split_stride=0, out_T=None, out_t_mul=1,
         out_t_add=0, bias=None, rowscale=None, out2=None, outp=None, outp_col0=0, engine=None, dbg=(0, 0, 0)):
    """Fill a drn_gemm_t.  a, b: L.Planes descriptors.  taps: sequence of (shift, parity, weight_tap).  See include/drn_b200.h."""
    g = L.GemmDesc()
    g.form, g.b_mn = form, b_mn
    g.a, g.b = a, b
    g.B, g.T, g.N, g.K, g.M = B, T, N, K, M
    g.ntaps = len(taps)
    for i, (sh, par, w) in enumerate(taps):
        g.tap_shift[i], g.tap_par[i], g.tap_w[i] = sh, par, w
    g.a_c0, g.b_c0 = a_c0, b_c0
    g.nprod = NPROD if nprod is None else nprod
    g.split_k = split_k
    if out is not None:
        g.out = out.data_ptr()
        g.out_ld = out_ld if out_ld is not None else out.shape[-1]
    g.out_col0, g.out_mode, g.out_tap_stride, g.out_split_stride = out_col0, out_mode, out_tap_stride, out_split_stride
    g.out_T = T if out_T is None else out_T
    g.out_t_mul, g.out_t_add = out_t_mul, out_t_add
    if bias is not None:
        g.bias = bias.data_ptr()
    if rowscale is not None:
        g.rowscale = rowscale.data_ptr()
        g.rowscale_ld = rowscale.shape[-1]
    if out2 is not None:
        g.out2 = out2.data_ptr()
        g.out2_ld = out2.shape[-1]
    if outp is not None:
        g.outp = outp.data.data_ptr()
        g.outp_ld = outp.C
        g.outp_col0 = outp_col0
        g.outp_plane_stride = outp.plane_stride
    g.engine = ENGINE if engine is None else engine
    g.dbg_lbo, g.dbg_sbo, g.dbg_kadv = dbg
    return g


# Schedule of the persistent contraction kernel for the launches that get a workspace (include/drn_b200.h:
# drn_gemm_set_schedule).  DRN_SCHEDULE = hybrid (default) | static | streamk.
#   hybrid : full waves on the static tile round-robin, the k-iterations of the last, partial wave cut into equal ranges over
#            all SM pairs and folded through the workspace (prop_fc weight gradient: 256 tiles = 3.46 waves on 74 pairs).
#   streamk: every tile boundary ignored.  Measured on B200 (r02, profiles/r02_ab_streamk.log): 3.86 ms per step against 3.45 ms
#            static -- contiguous per-pair tile ranges lose the L2 sharing of operand tiles between neighbouring SM pairs (prop_fc
#            forward 514 -> 624 us) and every extra segment pays a full TMEM -> global epilogue (conv1 backward 48 -> 79 us).
SCHEDULE = os.environ.get("DRN_SCHEDULE", "streamk" if os.environ.get("DRN_STREAMK", "0") == "1" else "hybrid")
if SCHEDULE not in ("static", "hybrid", "streamk"):
    raise ValueError("DRN_SCHEDULE must be static, hybrid or streamk")
_MODE = {"static": 0, "hybrid": 1, "streamk": 2}
STREAMK = SCHEDULE == "streamk"
_WS = {}


def workspace(device=None):
    """Workspace of the persistent contraction kernel (drn_gemm_workspace_bytes, include/drn_b200.h): one per device,
    zero-filled once, owned by the caller as every other buffer.  All contractions of a process run on one stream at a time
    (model/main_model.py), so one workspace per device is enough."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    ws = _WS.get(key)
    if ws is None:
        ws = torch.zeros(int(L.load().drn_gemm_workspace_bytes()), dtype=torch.uint8, device=torch.device("cuda", key))
        _WS[key] = ws
    return ws


def _ws_args(streamk):
    """streamk: None = the process schedule (DRN_SCHEDULE), False = static for this call (no workspace), True = full stream-K
    for this call (set by `_schedule`)."""
    if streamk is False or (streamk is None and SCHEDULE == "static"):
        return None, C.c_size_t(0)
    ws = workspace()
    return C.c_void_p(ws.data_ptr()), C.c_size_t(ws.numel())


class _schedule:
    """Context: full stream-K for the launches inside (tests, A/B), the process default restored afterwards."""

    def __init__(self, streamk):
        self.on = streamk is True

    def __enter__(self):
        if self.on:
            L.load().drn_gemm_set_schedule(2)

    def __exit__(self, *a):
        if self.on:
            L.load().drn_gemm_set_schedule(_MODE[SCHEDULE])


def gemm(*a, streamk=None, **k):
    """One contraction, one launch (engine chosen by the library)."""
    g = desc(*a, **k)
    ws, nb = _ws_args(streamk)
    with _schedule(streamk):
        L.check(L.load().drn_gemm_ws(C.byref(g), ws, nb, L.stream_ptr()), "drn_gemm")


GROUP_MAX = 6


def gemm_group(descs, streamk=None):
    """Independent contractions in ONE launch of the persistent CTA-pair kernel (drn_gemm_group).  Returns the launch count."""
    n = 0
    ws, nb = _ws_args(streamk)
    for i in range(0, len(descs), GROUP_MAX):
        chunk = descs[i:i + GROUP_MAX]
        arr = (L.GemmDesc * len(chunk))(*chunk)
        with _schedule(streamk):
            L.check(L.load().drn_gemm_group_ws(len(chunk), arr, ws, nb, L.stream_ptr()), "drn_gemm_group")
        n += 1
    return n
