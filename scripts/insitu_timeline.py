"""In-situ (warm caches, back to back, replayed from the CUDA graphs) duration of every library call of one training step.

ncu's launch list serialises kernels and flushes caches, which overstates the small HBM-bound kernels whose inputs are
L2-resident inside a real step; host events around eager launches are host-bound for those same kernels (measured, r01 v15:
4.9 ms eager against 3.7 ms replayed).  Here `drn_b200.lib.check` -- the single choke point every C-ABI call goes through --
is wrapped so that a one-thread `drn_timestamp` kernel (the GPU's %globaltimer) follows every call.  The stamps are captured
into the two CUDA graphs together with the kernels; after the warm-up the trace buffer is cleared, ONE step is replayed and the
stamps are sorted by time: the interval that ends at a stamp is the GPU time of the call before it (one call = one kernel,
except the query encoder's multi-kernel calls), plus the stamp kernel itself (~2 us, reported as `stamp_overhead_us` from
back-to-back stamps and subtracted).

    python scripts/insitu_timeline.py [--world-note ...]
Prints one JSON object: per-call-name totals and the sequence."""
import collections
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from drn_b200 import lib as L  # noqa: E402
from drn_b200 import spec as spec_mod  # noqa: E402
from drn_b200 import synthetic as S  # noqa: E402


def main():
    from model.main_model import mainModel
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    cfg = S.default_config(stage=1)
    sd = S.synth_state_dict(spec_mod.state_dict_spec(cfg))
    batch = S.synth_batch(32, 256, max_len=10, embedding=sd["query_encoder.embedding.weight"])
    model = mainModel(1301, S.config_namespace(stage=1))
    model.load_state_dict(sd)
    for k, prm in model.named_parameters():
        if "iou_scores" in k or "mix_fc" in k:
            prm.requires_grad = False
    model = model.to(dev).train()
    b = {k: v.to(dev) for k, v in batch.items()}
    b["query_length"] = batch["query_length"].to(dev)

    lib = L.load()
    trace = torch.zeros(8192, dtype=torch.int64, device=dev)
    names = []
    orig = L.check

    def stamp(what):
        i = len(names)
        if i >= trace.numel():
            return
        names.append(what)
        orig(lib.drn_timestamp(C.c_void_p(trace.data_ptr() + 8 * i), L.stream_ptr()), "timestamp")

    def check(rc, what=""):
        orig(rc, what)
        stamp(what)
    L.check = check

    def step():
        for prm in model.parameters():
            prm.grad = None
        stamp("(step begin)")
        _, ld = model(b["query_tokens"], b["query_length"], b["props_features"], b["props_start_end"], b["gt_start_end"], None, None)
        (ld["loss_cls"] + ld["loss_reg"] + ld["loss_iou"]).backward()
        stamp("(step end: autograd glue)")

    for _ in range(4):  # eager pass + graph captures + replays
        step()
    torch.cuda.synchronize()
    # stamp overhead: 16 back-to-back stamps
    o0 = len(names)
    for _ in range(16):
        stamp("(calibration)")
    torch.cuda.synchronize()
    cal = trace[o0:o0 + 16].cpu().tolist()
    overhead = sorted(cal[i + 1] - cal[i] for i in range(15))[7] / 1e3
    trace.zero_()
    step()  # ONE replayed step: the stamps captured in the graphs + the eager staging calls
    torch.cuda.synchronize()
    ts = trace.cpu().tolist()
    ev = sorted((t, names[i]) for i, t in enumerate(ts[:len(names)]) if t > 0)
    seq, agg = [], collections.defaultdict(float)
    for (t0, _), (t1, what) in zip(ev, ev[1:]):
        us = (t1 - t0) / 1e3
        seq.append((what, round(us, 1)))
        agg[what] += max(us - overhead, 0.0)
    total = (ev[-1][0] - ev[0][0]) / 1e3
    print(json.dumps({
        "note": "one training step replayed from the CUDA graphs with a %globaltimer stamp after every C-ABI call; microseconds; "
                "by_call_us has the stamp overhead subtracted, sequence_us has not",
        "stamp_overhead_us": round(overhead, 2), "stamps": len(ev), "step_us_with_stamps": round(total, 1),
        "by_call_us": {k: round(v, 1) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])},
        "sequence_us": seq}))


if __name__ == "__main__":
    main()
