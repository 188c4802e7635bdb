"""Bring-up probe for drn_gemm on a real B200 (run under gpurun).  Prints the error of every operand form for both
engines and, for the MN-major descriptor fields, sweeps candidate encodings.  Never asserts: one call = maximum info."""
import itertools
import os
import sys
import time
import traceback

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import test_gemm_gpu as G  # noqa: E402
from drn_b200 import lib as L  # noqa: E402
from drn_b200 import ops  # noqa: E402
from drn_b200.planes import Planes  # noqa: E402


def attempt(name, fn):
    try:
        err = fn()
        print("%-60s err=%.3e %s" % (name, err, "OK" if err < 2e-5 else "BAD"), flush=True)
        return err
    except Exception as e:  # noqa: BLE001
        print("%-60s EXC %s" % (name, str(e)[:200]), flush=True)
        traceback.print_exc()
        return float("inf")


def main():
    print("device check:", L.load().drn_device_check(), torch.cuda.get_device_name(0), flush=True)
    for eng in (1, 0):
        attempt("rows lin   eng%d" % eng, lambda: G.run_rows(2, 256, 128, 128, ((0, 0, 0),), engine=eng)[0])
        attempt("rows k3    eng%d" % eng, lambda: G.run_rows(2, 256, 256, 256, G.K3, engine=eng)[0])
        attempt("rows k3 T64 eng%d" % eng, lambda: G.run_rows(5, 64, 128, 256, G.K3, engine=eng)[0])
        attempt("rows k3s2  eng%d" % eng, lambda: G.run_rows(2, 128, 128, 256, G.K3S2, P=2, engine=eng)[0])
        e1 = attempt("rows dgrad (B MN-major) eng%d" % eng, lambda: G.run_rows(2, 256, 256, 128, G.K3, b_mn=1, engine=eng)[0])
        e2 = attempt("wgrad (A,B MN-major) eng%d" % eng, lambda: G.run_wgrad(2, 128, 256, 320, G.K3, engine=eng))
        if eng == 0 and (e1 > 2e-5 or e2 > 2e-5):
            print("--- sweeping MN-major descriptor encodings (lbo, sbo, kadv)", flush=True)
            for lbo, sbo, kadv in itertools.product((8192, 1024, 128, 16384), (1024, 8192, 128), (2048, 32, 256, 4096)):
                if lbo == sbo:
                    continue
                a = attempt("  dgrad lbo=%d sbo=%d kadv=%d" % (lbo, sbo, kadv),
                            lambda: G.run_rows(2, 256, 256, 128, G.K3, b_mn=1, engine=0, dbg=(lbo, sbo, kadv))[0])
                if a < 2e-5:
                    attempt("  wgrad same", lambda: G.run_wgrad(2, 128, 256, 320, G.K3, engine=0, dbg=(lbo, sbo, kadv)))
    # first timing of the dominant contraction: prop_fc forward, M=8192, N=K=4096
    try:
        B, T, D = 32, 256, 4096
        a = Planes.from_float(torch.randn(B, T, D, device="cuda").relu())
        w = Planes.from_float(torch.randn(1, D, D, device="cuda") / 64)
        out = torch.empty(B, T, D, device="cuda")
        for nprod in (3, 1):
            for _ in range(3):
                ops.gemm(L.GEMM_ROWS, a.desc(), w.desc(), B, T, D, K=D, out=out, nprod=nprod)
            torch.cuda.synchronize()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            ev[0].record()
            for _ in range(10):
                ops.gemm(L.GEMM_ROWS, a.desc(), w.desc(), B, T, D, K=D, out=out, nprod=nprod)
            ev[1].record()
            torch.cuda.synchronize()
            ms = ev[0].elapsed_time(ev[1]) / 10
            print("prop_fc fwd nprod=%d: %.3f ms  %.1f TFLOP/s algorithmic" % (nprod, ms, 2 * 8192 * 4096 * 4096 / ms / 1e9), flush=True)
        ref = (a.to_float()[0, :64].double() @ w.to_float()[0].double().t())
        ops.gemm(L.GEMM_ROWS, a.desc(), w.desc(), B, T, D, K=D, out=out, nprod=3)
        torch.cuda.synchronize()
        print("prop_fc check err %.3e" % ((out[0, :64].double() - ref).abs().max() / ref.abs().max()).item(), flush=True)
    except Exception:  # noqa: BLE001
        traceback.print_exc()


if __name__ == "__main__":
    main()
