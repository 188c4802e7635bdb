"""How long is the gap between two DEPENDENT kernel nodes of a CUDA graph on this GPU?  A step of the path is ~105 launches
replayed from graphs; this measures the floor every one of them pays: a graph of N one-thread kernels (drn_timestamp) in one
stream, replayed; time per node = launch gap + ~1 us of kernel.  Also the same chain behind a full-GPU persistent launch.
    python scripts/launch_gap_probe.py            (DRN_PDL=1 for programmatic dependent launch)"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from drn_b200 import lib as L  # noqa: E402


def main():
    lib = L.load()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    slots = torch.zeros(4096, dtype=torch.int64, device=dev)
    N = 400

    def chain():
        for i in range(N):
            L.check(lib.drn_timestamp(C.c_void_p(slots.data_ptr() + 8 * i), L.stream_ptr()), "ts")
    chain()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        chain()
    res = {}
    for rep in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    res["us_per_node_graph_chain"] = e0.elapsed_time(e1) * 1e3 / (5 * N)
    st = slots[:N].cpu().tolist()
    d = sorted(st[i + 1] - st[i] for i in range(N - 1))
    res["globaltimer_delta_ns_median"] = d[len(d) // 2]
    res["globaltimer_delta_ns_p10_p90"] = [d[len(d) // 10], d[9 * len(d) // 10]]
    e0.record()
    for _ in range(5):
        chain()
    e1.record()
    torch.cuda.synchronize()
    res["us_per_node_eager_chain"] = e0.elapsed_time(e1) * 1e3 / (5 * N)
    res["pdl"] = os.environ.get("DRN_PDL", "0")
    print(json.dumps(res))


if __name__ == "__main__":
    main()
