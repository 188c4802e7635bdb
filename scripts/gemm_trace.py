"""Where does a launch of the persistent contraction kernel spend its time?  In-kernel %globaltimer stamps (drn_gemm_trace) of
the 16 launches of one training step, replayed from the CUDA graphs (warm caches, back to back).  Per launch, microseconds
since the first CTA entered the kernel: entry of the last CTA, prologue done, first operands landed, last MMA issued, last
accumulator complete, last accumulator drained, exit.        python scripts/gemm_trace.py"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from drn_b200 import lib as L  # noqa: E402
from drn_b200 import spec as spec_mod  # noqa: E402
from drn_b200 import synthetic as S  # noqa: E402

NAMES = ["QE input projection", "prop_fc fwd (+gate)", "conv0 fwd (K-split 2)", "FPN inner x3 fwd", "FPN layer x3 fwd", "towers x3 fwd",
         "mix_fc x3 fwd", "iou_scores.0 x3 fwd", "towers bwd (dgrad+wgrad)", "FPN layer bwd", "FPN inner bwd", "conv2 bwd", "conv1 bwd",
         "conv0 bwd", "prop_fc wgrad", "QE bwd projections"]


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
    lib = L.load()
    CT = 160
    trace = torch.zeros(64 * CT * 8, dtype=torch.int64, device=dev)
    lib.drn_gemm_trace(C.c_void_p(trace.data_ptr()), 32)

    def step():
        for prm in model.parameters():
            prm.grad = None
        _, ld = model(b["query_tokens"], b["query_length"], b["props_features"], b["props_start_end"], b["gt_start_end"], None, None)
        (ld["loss_cls"] + ld["loss_reg"] + ld["loss_iou"]).backward()

    for _ in range(4):  # eager + capture (slots 0-7 / 8-15 forward, 16-23 / 24-31 backward), then replays
        step()
    torch.cuda.synchronize()
    trace.zero_()
    step()
    torch.cuda.synchronize()
    t = trace.view(64, CT, 8).cpu()
    out = []
    slots = list(range(8, 16)) + list(range(24, 32))
    prev_exit = None
    for name, sl in zip(NAMES, slots):
        ctas, tiles, lpt = C.c_int(0), C.c_int(0), C.c_int(0)
        lib.drn_gemm_trace_info(sl, C.byref(ctas), C.byref(tiles), C.byref(lpt))
        n = ctas.value
        x = t[sl, :n].double()
        t0 = float(x[:, 0].min())
        us = lambda v: round((float(v) - t0) / 1e3, 2)  # noqa: E731
        lead = x[0::2]  # leader CTAs own the MMA stamps
        row = {"launch": name, "ctas": n, "tiles": tiles.value, "balanced_schedule": lpt.value,
               "last_cta_entered": us(x[:, 0].max()), "prologue_done_median": us(x[:, 1].median()),
               "first_operands_landed_median": us(lead[:, 2].median()), "last_mma_issued_max": us(lead[:, 3].max()),
               "last_mma_issued_median": us(lead[:, 3].median()),
               "last_accumulator_complete_max": us(torch.maximum(x[:, 4], x[:, 5]).max()),
               "drained_max": us(x[:, 6].max()), "drained_median": us(x[:, 6].median()), "exit_max": us(x[:, 7].max()),
               "since_previous_traced_exit": None if prev_exit is None else round((t0 - prev_exit) / 1e3, 2)}
        prev_exit = float(x[:, 7].max())
        out.append(row)
    print(json.dumps({"note": "in-kernel stamps of the 16 pair-kernel launches of one replayed step; us since the first CTA's entry",
                      "launches": out}))


if __name__ == "__main__":
    main()
