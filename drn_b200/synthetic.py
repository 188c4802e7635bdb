"""Seeded synthetic configs, weights and batches for the DRN dense-regression hot path.

Everything here is generated with numpy's PCG64 (bit-stable across machines), never with
torch's RNG, so the golden-vector generator (oracle/make_goldens.py, run where
/root/reference exists) and the tests / bench (run on a GPU box where it does not) see
byte-identical weights and inputs.

Input recipe: SURVEY.md section 8d.  Config keys: reference data/default_config.yaml:2-66.
"""
import zlib
from argparse import Namespace

import numpy as np
import torch

SEED = 222  # the reference's own (commented-out) seed, main.py:35


def default_config(stage=1, feature_type="C3D", **overrides):
    """The `Charades` block of the reference's data/default_config.yaml merged with the
    stage flags main.py gets from opts.py:11-13.  Returned as a plain dict."""
    cfg = {
        "MFnet": {"feature_dim": 768, "ft_window_size": 32, "ft_overlap": 0.75},
        "C3D": {"feature_dim": 4096, "ft_window_size": 16, "ft_overlap": 0.5},
        "I3D": {"feature_dim": 2048, "ft_window_size": 64, "ft_overlap": 0.875},
        "dataset": "Charades",
        "feature_type": feature_type,
        "n_epoch": 50,
        "batch_size": 32,
        "test_batch_size": 16,
        "lr": 0.001,
        "loss_weights": 0.5,
        "clip_gradient": 0.5,
        "loss_type": "iou",
        "lstm_layers": 1,
        "hidden_dim": 512,
        "embedding": 300,
        "node_ft_dim": 1024,
        "graph_num_layers": 3,
        "start_epoch": 0,
        "weight_decay": 5e-4,
        "pos_thr": 0.5,
        "neg_thr": 0.1,
        "first_output_dim": 256,
        "fpn_feature_dim": 512,
        "fpn_stride": [1, 2, 4],
        "fcos_conv_layers": 1,
        "fcos_prior_prob": 0.01,
        "fcos_loss_alpha": 0.25,
        "fcos_loss_gamma": 2.0,
        "fcos_inference_thr": 0.05,
        "fcos_pre_nms_top_n": 32,
        "fcos_nms_thr": 0.6,
        "fcos_num_class": 2,
        "test_detections_per_img": 32,
        "is_first_stage": stage == 1,
        "is_second_stage": stage == 2,
        "is_third_stage": stage == 3,
    }
    cfg.update(overrides)
    return cfg


def config_namespace(stage=1, **overrides):
    """mainModel takes an argparse.Namespace and calls vars() on it (main_model.py:17)."""
    return Namespace(**default_config(stage=stage, **overrides))


def _rng(seed, name):
    return np.random.Generator(np.random.PCG64([int(seed), zlib.crc32(name.encode())]))


_XAVIER_KEYS = ("query_encoder.qInput", "query_encoder.cmd_inter2logits")


def synth_tensor(name, shape, seed=SEED):
    """Deterministic value for one state_dict entry, scaled like the reference initialiser of
    that entry (fcos.py:72-85, basic_blocks.py:20, ops.py:20-25, torch defaults) but with
    non-trivial biases / BN affine / running stats so every term of the math is exercised."""
    g = _rng(seed, name)
    shape = tuple(shape)

    def normal(std, mean=0.0):
        return (g.standard_normal(shape, dtype=np.float32) * np.float32(std) + np.float32(mean)).astype(np.float32)

    def uniform(bound):
        return ((g.random(shape, dtype=np.float32) * 2 - 1) * np.float32(bound)).astype(np.float32)

    if name.endswith("num_batches_tracked"):
        return torch.zeros(shape, dtype=torch.int64)
    if name.endswith("running_mean"):
        return torch.from_numpy(normal(0.1))
    if name.endswith("running_var"):
        return torch.from_numpy((g.random(shape, dtype=np.float32) + np.float32(0.5)).astype(np.float32))
    if name.endswith(".scale"):
        return torch.from_numpy(normal(0.05, 1.0))
    if name.endswith("embedding.weight"):
        w = normal(0.35)
        w[0] = 0  # padding_idx = 0 (language_module.py:14)
        return torch.from_numpy(w)
    if "biLSTM" in name:
        return torch.from_numpy(uniform(1.0 / np.sqrt(512.0)))
    if len(shape) == 1:
        if name.endswith("weight"):  # BatchNorm gamma
            return torch.from_numpy(normal(0.1, 1.0))
        if name.endswith("cls_logits.bias"):
            return torch.from_numpy(normal(0.05, -4.59512))  # -log(99), fcos.py:81-83
        return torch.from_numpy(normal(0.05))
    if len(shape) == 3:
        if "fcos.head" in name:
            return torch.from_numpy(normal(0.01))
        fan_in = shape[1] * shape[2]
        return torch.from_numpy(uniform(np.sqrt(3.0 / fan_in)))
    if len(shape) == 2:
        if any(k in name for k in _XAVIER_KEYS):
            return torch.from_numpy(uniform(np.sqrt(6.0 / (shape[0] + shape[1]))))
        return torch.from_numpy(uniform(1.0 / np.sqrt(shape[1])))
    raise ValueError("no rule for %s %s" % (name, shape))


_CHARADES = None


def charades_fixture():
    """tests/golden/charades_queries.npz (written by oracle/make_query_fixture.py from the data files the reference ships):
    real Charades-STA queries tokenised through Charades_word2id.json and the GloVe-300 embedding table of data/glove_weights
    (SURVEY.md section 8d).  Used by bench.py / tests / scripts for their inputs; never by the product path."""
    global _CHARADES
    if _CHARADES is None:
        import os
        path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "charades_queries.npz")
        z = np.load(path)
        _CHARADES = {k: z[k] for k in z.files}
    return _CHARADES


def synth_state_dict(spec, seed=SEED, glove=False):
    """spec: iterable of (name, shape) (e.g. from model.state_dict()).  Returns name->tensor.  glove=True: the embedding table is
    the reference's data/glove_weights (what main.py:94 copies in), not a random stand-in."""
    sd = {name: synth_tensor(name, shape, seed) for name, shape in spec}
    if glove:
        g = torch.from_numpy(charades_fixture()["glove"].copy())
        assert tuple(sd["query_encoder.embedding.weight"].shape) == tuple(g.shape)
        sd["query_encoder.embedding.weight"] = g
    return sd


def load_synth_weights(model, seed=SEED):
    """Overwrite every entry of `model.state_dict()` with its synthetic value."""
    spec = [(k, tuple(v.shape)) for k, v in model.state_dict().items()]
    sd = synth_state_dict(spec, seed)
    model.load_state_dict(sd)
    return sd


def synth_batch(B, T, max_len=10, feature_dim=4096, vocab_size=1301, seed=SEED, embedding=None,
                sorted_lengths=True, queries="random", split="train", signal_scale=2.0):
    """One synthetic batch in the dtypes dataset.py:180-224 produces.  queries="charades": real Charades-STA queries drawn from
    the fixture (`split` = train | test for held-out evaluation), padded to the longest of the batch like collate_data.

    returns dict(query_tokens i64 [B,L], query_length i64 [B] (descending), props_features f32
    [B,T,D], props_start_end f64 [B,T,2], gt_start_end f64 [B,2])."""
    g = _rng(seed, "batch/%d/%d/%d" % (B, T, max_len))
    if queries == "charades":
        fx = charades_fixture()
        pool_t, pool_l = fx[split + "_tokens"], fx[split + "_len"]
        ok = np.nonzero(pool_l <= max_len)[0]
        pick = ok[g.integers(0, len(ok), size=B)]
        lengths = pool_l[pick].astype(np.int64)
        order = np.argsort(-lengths, kind="stable") if sorted_lengths else np.arange(B)  # dataset.py:183
        pick, lengths = pick[order], lengths[order]
        max_len = int(lengths.max())  # collate_data pads to the longest query of the batch (dataset.py:186)
        tokens = pool_t[pick, :max_len].astype(np.int64)
    else:
        lengths = g.integers(2, max_len + 1, size=B)
        lengths[0] = max_len
        if sorted_lengths:
            lengths = np.sort(lengths)[::-1].copy()
        tokens = np.zeros((B, max_len), dtype=np.int64)
        for b in range(B):
            tokens[b, : lengths[b]] = g.integers(1, vocab_size + 1, size=lengths[b])
    c = g.uniform(0.2, 0.8, size=B)
    w = g.uniform(0.05, 0.35, size=B)
    gt = np.stack([np.clip(c - w, 0.0, 1.0), np.clip(c + w, 0.0, 1.0)], axis=1)  # f64
    t = np.arange(T, dtype=np.float64)
    pse = np.broadcast_to(np.stack([t / T, (t + 1) / T], axis=1)[None], (B, T, 2)).copy()
    feats = np.maximum(g.standard_normal((B, T, feature_dim), dtype=np.float32), 0)
    if embedding is not None:
        emb = embedding.detach().cpu().numpy().astype(np.float32)
        proj = _rng(seed, "signal_proj").standard_normal((emb.shape[1], feature_dim), dtype=np.float32)
        proj /= np.float32(np.sqrt(emb.shape[1]))
        loc = np.arange(T, dtype=np.float32) + 0.5
        for b in range(B):
            q = emb[tokens[b, : lengths[b]]].mean(0)
            sig = np.maximum(q @ proj, 0) * np.float32(signal_scale)  # 2.0: SURVEY 8d; smaller = a harder task (R@1 parity runs)
            inside = (loc > 32 * gt[b, 0]) & (loc < 32 * gt[b, 1])
            feats[b, inside] += sig
    return {
        "query_tokens": torch.from_numpy(tokens),
        "query_length": torch.from_numpy(lengths.astype(np.int64)),
        "props_features": torch.from_numpy(feats),
        "props_start_end": torch.from_numpy(pse),
        "gt_start_end": torch.from_numpy(gt),
    }


def craft_stage23(sd, batch):
    """Make the IoU-score branch (loss.py:168-198) active: with random weights no prediction ever has
    tIoU > 0.9 (hard part 4), so shrink bbox_pred to ~constant exp(bias) = 5 location
    units and give every sample a GT of half-width 5/32 centred on a level-0 location."""
    sd = dict(sd)
    sd["fcos.head.bbox_pred.weight"] = sd["fcos.head.bbox_pred.weight"] * 0.01
    sd["fcos.head.bbox_pred.bias"] = torch.full_like(sd["fcos.head.bbox_pred.bias"], float(np.log(5.0)))
    for l in range(3):
        sd["fcos.head.scales.%d.scale" % l] = torch.ones_like(sd["fcos.head.scales.%d.scale" % l])
    batch = dict(batch)
    B = batch["gt_start_end"].shape[0]
    centre = (6.0 + 2.0 * np.arange(B, dtype=np.float64) + 0.5 + 0.05) / 32.0
    batch["gt_start_end"] = torch.from_numpy(np.stack([centre - 5.0 / 32, centre + 5.0 / 32], axis=1))
    return sd, batch


GOLDEN_CASES = {
    # name: (B, T, max_len, stage, training, crafted)
    "c1_eval_b1_t64": (1, 64, 10, 1, False, False),
    "c1_train_b1_t64": (1, 64, 10, 1, True, False),
    "s1_train_b4_t32": (4, 32, 8, 1, True, False),
    "s1_train_b2_t256": (2, 256, 10, 1, True, False),
    "s3_train_b4_t32_crafted": (4, 32, 7, 3, True, True),
    "s2_train_b4_t32_crafted": (4, 32, 7, 2, True, True),
    "s3_eval_b3_t64_crafted": (3, 64, 9, 3, False, True),
    # real Charades-STA queries + the GloVe-300 table (tests/golden/charades_queries.npz), SURVEY.md section 8d
    "s1_train_b4_t32_charades": (4, 32, 10, 1, True, False),
}


def golden_case(name, spec):
    """Rebuild (cfg, state_dict, batch) of a golden case from the seed alone."""
    B, T, L, stage, training, crafted = GOLDEN_CASES[name]
    cfg = default_config(stage=stage)
    real = name.endswith("_charades")
    sd = synth_state_dict(spec, glove=real)
    batch = synth_batch(B, T, max_len=L, embedding=sd["query_encoder.embedding.weight"], queries="charades" if real else "random")
    if crafted:
        sd, batch = craft_stage23(sd, batch)
    return cfg, sd, batch, stage, training


def sample_indices(numel, n=16):
    """Deterministic flat indices used to spot-check big tensors in the golden files."""
    if numel <= n:
        return np.arange(numel)
    return (np.arange(n, dtype=np.int64) * 2654435761 % numel).astype(np.int64)
