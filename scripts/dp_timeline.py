"""Timeline of ONE data-parallel training step per rank (run under torchrun): CUDA events recorded by model/main_model.py's
DP_TRACE hooks between the phases of the backward (graph 1 | tail | prop_fc wgrad chunks | the points at which each all-reduce
had completed), after warm-up, averaged over `--steps` steps.  Rank 0 prints one JSON object with every rank's phase times (ms
since the start of the backward): which collective is exposed, and how much the compute phases stretch beside NCCL
(compare with the same phases at world size 1).   torchrun --nproc-per-node 8 scripts/dp_timeline.py [--steps 20]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from drn_b200 import spec as spec_mod  # noqa: E402
from drn_b200 import synthetic as S  # noqa: E402
from drn_b200.parallel import DataParallelDRN  # noqa: E402
from model import main_model as MM  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        from drn_b200.parallel import nccl_env_defaults
        if os.environ.get("DRN_NCCL_MAX_CTAS") != "default":
            nccl_env_defaults()
        dist.init_process_group("nccl", device_id=dev)
    cfg = S.default_config(stage=1)
    sd = S.synth_state_dict(spec_mod.state_dict_spec(cfg), glove=True)
    batch = S.synth_batch(32, 256, max_len=10, embedding=sd["query_encoder.embedding.weight"], seed=S.SEED + rank, queries="charades")
    model = MM.mainModel(1301, S.config_namespace(stage=1))
    model.load_state_dict(sd)
    for k, p in model.named_parameters():
        if "iou_scores" in k or "mix_fc" in k:
            p.requires_grad = False
    model = model.to(dev).train()
    if world > 1:
        model = DataParallelDRN(model)
    core = model.module if world > 1 else model
    b = {k: v.to(dev) for k, v in batch.items()}

    def step():
        for p in core.parameters():
            p.grad = None
        _, ld = model(b["query_tokens"], b["query_length"], b["props_features"], b["props_start_end"], b["gt_start_end"], None, None)
        (ld["loss_cls"] + ld["loss_reg"] + ld["loss_iou"]).backward()
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    acc, names = None, None
    for _ in range(a.steps):
        if world > 1:
            dist.barrier()
        MM.DP_TRACE = []
        e0 = torch.cuda.Event(enable_timing=True)
        e0.record()
        step()
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record()
        torch.cuda.synchronize()
        tr = MM.DP_TRACE
        MM.DP_TRACE = None
        t = [e0.elapsed_time(tr[0][1])] + [tr[0][1].elapsed_time(e) for _, e in tr[1:]] + [tr[0][1].elapsed_time(e1)]
        names = ["forward (step start -> backward start)"] + [n for n, _ in tr[1:]] + ["step end"]
        acc = t if acc is None else [x + y for x, y in zip(acc, t)]
    mine = [x / a.steps for x in acc]
    if world > 1:
        allt = [None] * world
        dist.all_gather_object(allt, mine)
    else:
        allt = [mine]
    if rank == 0:
        print(json.dumps({"world": world, "order": os.environ.get("DRN_DP_ORDER", "tail_first"), "chunks": os.environ.get("DRN_DP_CHUNKS", "4"),
                          "nccl_max_ctas": os.environ.get("NCCL_MAX_CTAS"), "steps": a.steps,
                          "unit": "ms since the start of the backward (first entry: forward duration)", "phases": names,
                          "per_rank": [[round(x, 4) for x in r] for r in allt],
                          "max_over_ranks": [round(max(r[i] for r in allt), 4) for i in range(len(names))]}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
