"""Bring-up probe of the persistent CTA-pair kernel (engine 2) on a B200: errors per operand form + timing of the big shapes."""
import os
import sys
import traceback

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gemm_gpu as G  # noqa: E402
from drn_b200 import lib as L  # noqa: E402
from drn_b200 import ops  # noqa: E402
from drn_b200.planes import Planes  # noqa: E402


def attempt(name, fn):
    try:
        err = fn()
        print("%-50s err=%.3e %s" % (name, err, "OK" if err < 2e-5 else "BAD"), flush=True)
    except Exception as e:  # noqa: BLE001
        print("%-50s EXC %s" % (name, str(e)[:300]), flush=True)
        traceback.print_exc()


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    eng = 2
    attempt("rows lin", lambda: G.run_rows(2, 256, 128, 256, ((0, 0, 0),), engine=eng)[0])
    attempt("rows k3", lambda: G.run_rows(2, 256, 256, 256, G.K3, engine=eng)[0])
    attempt("rows k3 T64 B5", lambda: G.run_rows(5, 64, 128, 256, G.K3, engine=eng)[0])
    attempt("rows k3 N1000 T96", lambda: G.run_rows(2, 96, 64, 1000, G.K3, engine=eng)[0])
    attempt("rows k3s2", lambda: G.run_rows(2, 128, 128, 256, G.K3S2, P=2, engine=eng)[0])
    attempt("rows dgrad", lambda: G.run_rows(2, 256, 256, 128, G.K3, b_mn=1, engine=eng)[0])
    attempt("rows many tiles", lambda: G.run_rows(16, 256, 128, 2048, G.K3, engine=eng)[0])
    attempt("wgrad", lambda: G.run_wgrad(2, 128, 256, 320, G.K3, engine=eng))
    attempt("wgrad split", lambda: G.run_wgrad(8, 128, 256, 256, G.K3, split_k=4, engine=eng))
    attempt("wgrad many", lambda: G.run_wgrad(16, 256, 1024, 1536, G.K3, engine=eng))
    B, T, D = 32, 256, 4096
    a = Planes.from_float(torch.randn(B, T, D, device="cuda").relu())
    w = Planes.from_float(torch.randn(1, D, D, device="cuda") / 64)
    out = torch.empty(B, T, D, device="cuda")
    dy = Planes.from_float(torch.randn(B, T, D, device="cuda"))
    for e in (3, 2):
        ms = timeit(lambda: ops.gemm(L.GEMM_ROWS, a.desc(), w.desc(), B, T, D, K=D, out=out, engine=e))
        print("prop_fc fwd   engine %d: %.3f ms  %.1f TFLOP/s algorithmic" % (e, ms, 2 * 8192 * 4096 * 4096 / ms / 1e9), flush=True)
        ms = timeit(lambda: ops.gemm(L.GEMM_WGRAD, dy.desc(), a.desc(), B, T, D, M=D, out=out.view(-1)[:D * D].view(D, D), out_ld=D, engine=e))
        print("prop_fc wgrad engine %d: %.3f ms  %.1f TFLOP/s algorithmic" % (e, ms, 2 * 8192 * 4096 * 4096 / ms / 1e9), flush=True)
    # conv0 shapes
    x0 = Planes.from_float(torch.randn(B, T, 4352, device="cuda"))
    w0 = Planes.from_float(torch.randn(3, 256, 4352, device="cuda") / 100)
    y0 = torch.empty(B, T, 256, device="cuda")
    for e in (3, 2):
        ms = timeit(lambda: ops.gemm(L.GEMM_ROWS, x0.desc(), w0.desc(), B, T, 256, K=4352, taps=G.K3, out=y0, engine=e))
        print("conv0 fwd     engine %d: %.3f ms  %.1f TFLOP/s algorithmic" % (e, ms, 2 * 8192 * 256 * 3 * 4352 / ms / 1e9), flush=True)
    # mid-size: FPN layer1 (M=8192, N=512, K=3*512)
    xl = Planes.from_float(torch.randn(B, T, 512, device="cuda"))
    wl = Planes.from_float(torch.randn(3, 512, 512, device="cuda") / 40)
    yl = torch.empty(B, T, 512, device="cuda")
    for e in (3, 2):
        ms = timeit(lambda: ops.gemm(L.GEMM_ROWS, xl.desc(), wl.desc(), B, T, 512, K=512, taps=G.K3, out=yl, engine=e))
        print("fpn layer1    engine %d: %.3f ms  %.1f TFLOP/s algorithmic" % (e, ms, 2 * 8192 * 512 * 3 * 512 / ms / 1e9), flush=True)


if __name__ == "__main__":
    main()
