"""TEST / MEASUREMENT INFRASTRUCTURE ONLY (never imported by the product path).

Times the UNMODIFIED reference (Alvin-Zeng/DRN `model.main_model.mainModel`, its own PyTorch code) on the host CPU cores:
forward + backward of BASELINE configs[1] (B=32, T=256, first stage) on the same seeded weights and batch as bench.py's own
arm.  The reference sources travel to the GPU box as ONE git-ignored archive, oracle/_ref/drn_reference.zip, which
`__graft_entry__.build()` packs from /root/reference in the build container (nothing of it enters the history); here it is
unpacked into a temporary directory and imported through oracle/ref_loader.py + oracle/shims.

Run as a subprocess of bench.py (the reference's package is called `model`, like the repo's drop-in package) with
CUDA_VISIBLE_DEVICES="" -- the reference calls `.cuda()` unconditionally (loss.py:239), which ref_loader turns into a no-op
only when no GPU is visible.  Prints one JSON object."""
import argparse
import json
import os
import sys
import tempfile
import time
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
ARCHIVE = os.path.join(HERE, "_ref", "drn_reference.zip")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--B", type=int, default=32)
    ap.add_argument("--T", type=int, default=256)
    a = ap.parse_args()
    if not os.path.isfile(ARCHIVE):
        print(json.dumps({"unavailable": "oracle/_ref/drn_reference.zip not built (run __graft_entry__.build() where /root/reference exists)"}))
        return
    tmp = tempfile.mkdtemp(prefix="drn_ref_")
    with zipfile.ZipFile(ARCHIVE) as z:
        z.extractall(tmp)
    os.environ["DRN_REFERENCE_ROOT"] = tmp
    os.chdir(tmp)
    sys.path.insert(0, REPO)
    import torch
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    from drn_b200 import synthetic as S
    from oracle import ref_loader
    cfg = S.default_config(stage=1)
    model = ref_loader.build_reference_model(cfg)
    sd = S.synth_state_dict([(k, tuple(v.shape)) for k, v in model.state_dict().items()], glove=True)  # as bench.build_inputs
    model.load_state_dict(sd)
    model.train()
    batch = S.synth_batch(a.B, a.T, max_len=10, embedding=sd["query_encoder.embedding.weight"], queries="charades")
    times, loss_v = [], None
    for i in range(a.warmup + a.steps):
        t0 = time.perf_counter()
        for p in model.parameters():
            p.grad = None
        _, ld = model(batch["query_tokens"], batch["query_length"], batch["props_features"], batch["props_start_end"],
                      batch["gt_start_end"], None, None)
        loss = sum(v for v in ld.values())
        loss.backward()
        loss_v = float(loss)
        if i >= a.warmup:
            times.append(time.perf_counter() - t0)
    print(json.dumps({"ms_per_step": 1e3 * sum(times) / len(times), "steps": a.steps, "warmup": a.warmup, "cores": cores,
                      "loss": loss_v, "B": a.B, "T": a.T, "torch": torch.__version__,
                      "impl": "unmodified reference model.main_model.mainModel (CPU, torch fp32)"}))


if __name__ == "__main__":
    main()
