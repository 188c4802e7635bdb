"""In-situ (warm-cache, back-to-back) duration of every library call of one training step.

ncu's launch list serialises kernels and flushes caches, which overstates the small HBM-bound kernels whose inputs are
L2-resident inside a real step.  Here the step runs eagerly (DRN_NO_GRAPHS=1) with a CUDA event recorded after every C-ABI
call (drn_b200.lib.check is the single choke point); the host stays ahead of the GPU, so the interval between two consecutive
events is the GPU time of the call in between (one call = one kernel, except the query encoder's two multi-kernel calls).

CAVEAT (measured, r01 v15): the eager step is host-bound around the small kernels (4.9 ms against 3.7 ms replayed from the
graphs: a ctypes call + event record costs ~10 us), so only the intervals of calls longer than ~20 us are GPU time; for the
small kernels use the ncu section captures (scripts/gpu_ncu_small.sh), which have the opposite bias (cold caches).

    DRN_NO_GRAPHS=1 python scripts/insitu_timeline.py [--steps 5]
Prints per-call medians over the steps, aggregated by call name, as one JSON object."""
import argparse
import collections
import json
import os
import statistics
import sys

os.environ["DRN_NO_GRAPHS"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from drn_b200 import lib as L  # noqa: E402
from drn_b200 import spec as spec_mod  # noqa: E402
from drn_b200 import synthetic as S  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    a = ap.parse_args()
    from model.main_model import mainModel
    dev = torch.device("cuda", 0)
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
    b["query_length"] = batch["query_length"]

    marks = []
    orig = L.check

    def check(rc, what=""):
        orig(rc, what)
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        marks.append((what, e))
    L.check = check
    import drn_b200.dense as D
    D.L.check = check

    def step():
        for prm in model.parameters():
            prm.grad = None
        _, ld = model(b["query_tokens"], b["query_length"], b["props_features"], b["props_start_end"], b["gt_start_end"], None, None)
        (ld["loss_cls"] + ld["loss_reg"] + ld["loss_iou"]).backward()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    per = collections.defaultdict(list)
    totals = []
    for _ in range(a.steps):
        marks.clear()
        e0 = torch.cuda.Event(enable_timing=True)
        e0.record()
        step()
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record()
        torch.cuda.synchronize()
        prev = e0
        seq = collections.Counter()
        for what, e in marks:
            seq[what] += 1
            per["%s#%d" % (what, seq[what])].append(prev.elapsed_time(e) * 1e3)
            prev = e
        per["(after last call: autograd glue)"].append(prev.elapsed_time(e1) * 1e3)
        totals.append(e0.elapsed_time(e1) * 1e3)
    rows = [(k, statistics.median(v)) for k, v in per.items()]
    agg = collections.defaultdict(float)
    for k, v in rows:
        agg[k.split("#")[0]] += v
    out = {"note": "eager step, event after every C-ABI call, medians over %d steps, microseconds" % a.steps,
           "step_us_eager": statistics.median(totals),
           "by_call_us": {k: round(v, 1) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])},
           "sequence_us": [(k, round(v, 1)) for k, v in rows]}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
