"""TEST INFRASTRUCTURE ONLY.  Writes tests/golden/pool_props.npz by running the UNMODIFIED reference dataset code
(/root/reference/dataset.py: CharadesSTA.get_data + collate_data) on seeded synthetic C3D-like feature files (D = 64 so the
full outputs fit in the repository; the code is dimension-agnostic), including videos with FEWER feature windows than their
frame count implies (the clamp of dataset.py:145).

    cd /root/reference && python /root/repo/oracle/make_pool_goldens.py       # needs /root/reference; CPU only
"""
import os
import sys
import tempfile
import types
from argparse import Namespace

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
nltk = types.ModuleType("nltk")  # dataset.py:8,92 -- only word_tokenize is used; whitespace split is enough for fixtures
nltk.word_tokenize = lambda s: s.split()
sys.modules["nltk"] = nltk
sys.path.insert(0, "/root/reference")
os.chdir("/root/reference")
sys.dont_write_bytecode = True
import dataset as ref_dataset  # noqa: E402

import yaml  # noqa: E402

cfg = yaml.safe_load(open("data/default_config.yaml"))["Charades"]
cfg["feature_type"] = "C3D"
D = 64
rng = np.random.default_rng(222)
with tempfile.TemporaryDirectory() as d:
    cfg["C3D"]["feature_root"] = d
    # word2id misses a few whitespace-split tokens (punctuation): map unknown words to id 1 for the fixture run
    import json
    w2i = json.load(open("data/dataset/Charades/Charades_word2id.json"))

    class _W(dict):
        def __missing__(self, k):
            return 1
    orig_load = json.load
    json.load = lambda f: _W(orig_load(f)) if "word2id" in getattr(f, "name", "") else orig_load(f)
    ds = ref_dataset.CharadesSTA(Namespace(**cfg), split="test")
    json.load = orig_load
    picks = [0, 7, 19, 101, 555, 1234]
    feats = {}
    for n, i in enumerate(picks):
        v = ds.video_list[i]
        full = max(1, (v.num_frames - 16) // 8 + 1)
        n_win = full if n % 2 == 0 else max(1, full // 2)  # every other video is short: exercises the clamp
        if v.id not in feats:
            feats[v.id] = rng.standard_normal((n_win, D)).astype(np.float32)
            torch.save(torch.from_numpy(feats[v.id]), os.path.join(d, v.id + ".pt"))
    batch = [ds[i] for i in picks]
    (vid_names, props_s_e, props_features, gt_start_end, query_tokens, query_length, props_num, num_frames) = ref_dataset.collate_data(batch)
    out = {"props_features": props_features.numpy(), "props_s_e": props_s_e.numpy(), "props_num": props_num.numpy(),
           "num_frames": num_frames.numpy(), "vid_names": np.array(vid_names)}
    for k, name in enumerate(vid_names):
        v = [ds.video_list[i] for i in picks if ds.video_list[i].id == name][0]
        out["feat_%d" % k] = feats[name]
        out["p_start_%d" % k] = np.array([p.start_frame for p in v.proposals], dtype=np.float64)
        out["p_end_%d" % k] = np.array([p.end_frame for p in v.proposals], dtype=np.int64)
    np.savez_compressed(os.path.join(REPO, "tests", "golden", "pool_props.npz"), **out)
    print("wrote", {k: getattr(v, "shape", None) for k, v in out.items() if not k.startswith(("feat_", "p_"))})
