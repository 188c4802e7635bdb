"""CPU-only checks: the C-ABI library loads and exports every symbol include/drn_b200.h declares, the ctypes mirrors match
the C structs, the drop-in model keeps the reference's state_dict contract and refuses to run without a GPU, the tap
tables of the host schedule describe the reference convolutions, and the host-side post-processing matches the oracle."""
import ctypes
import os
import re
import subprocess
import tempfile

import pytest
import torch
import torch.nn.functional as F

from drn_b200 import dense
from drn_b200 import lib as L
from drn_b200 import spec as spec_mod
from drn_b200 import synthetic as S

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(REPO, "include", "drn_b200.h")


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(drn_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = L.load()
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "libdrn_sm100.so does not export %s" % n
    assert lib.drn_version() == 100


def test_ctypes_structs_match_c_layout():
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "drn_b200.h"
int main(void) {
  printf("%zu %zu ", sizeof(drn_adam_item_t), offsetof(drn_adam_item_t, update));
  printf("%zu %zu %zu %zu ", sizeof(drn_sgemm_job_t), offsetof(drn_sgemm_job_t, bias), sizeof(drn_linear_job_t), offsetof(drn_linear_job_t, relu));
  printf("%zu %zu %zu ", sizeof(drn_head_levels_t), offsetof(drn_head_levels_t, tower), offsetof(drn_head_levels_t, d_tower));
  printf("%zu %zu %zu %zu ", sizeof(drn_bn_job_t), offsetof(drn_bn_job_t, coef), offsetof(drn_bn_job_t, out_qa), offsetof(drn_bn_job_t, dy_plane_stride));
  printf("%zu %zu %zu %zu ", sizeof(drn_qe_t), offsetof(drn_qe_t, tokens), offsetof(drn_qe_t, g_w2), offsetof(drn_qe_t, workspace_bytes));
  printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(drn_planes_t), sizeof(drn_gemm_t), offsetof(drn_gemm_t, b), offsetof(drn_gemm_t, tap_w),
         offsetof(drn_gemm_t, out_split_stride), offsetof(drn_gemm_t, outp_plane_stride), offsetof(drn_gemm_t, dbg_kadv),
         sizeof(drn_bn_part_t), offsetof(drn_bn_part_t, dbeta), sizeof(drn_pack_item_t), offsetof(drn_pack_item_t, slice_stride));
  printf("%zu %zu %zu ", offsetof(drn_gemm_t, stats), offsetof(drn_bn_job_t, partials), offsetof(drn_bn_job_t, partial_rows));
  printf("%zu %zu %zu %d %d\n", sizeof(drn_p2p_t), offsetof(drn_p2p_t, buf), offsetof(drn_p2p_t, flags), DRN_P2P_MAX_RANKS, DRN_P2P_FLAG_WORDS);
  return 0;
}'''
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c")
        open(c, "w").write(prog)
        exe = os.path.join(d, "t")
        subprocess.run(["gcc", "-I", os.path.join(REPO, "include"), c, "-o", exe], check=True)
        out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()
    G = L.GemmDesc
    J = L.BnJob
    HL = L.HeadLevels
    from drn_b200.optim import _Item
    mine = [ctypes.sizeof(_Item), _Item.update.offset, ctypes.sizeof(L.SgemmJob), L.SgemmJob.bias.offset, ctypes.sizeof(L.LinearJob), L.LinearJob.relu.offset, ctypes.sizeof(HL), HL.tower.offset, HL.d_tower.offset, ctypes.sizeof(J), J.coef.offset, J.out_qa.offset, J.dy_plane_stride.offset, ctypes.sizeof(L.Qe), L.Qe.tokens.offset, L.Qe.g_w2.offset, L.Qe.workspace_bytes.offset, ctypes.sizeof(L.Planes), ctypes.sizeof(G), G.b.offset, G.tap_w.offset, G.out_split_stride.offset,
            G.outp_plane_stride.offset, G.dbg_kadv.offset, ctypes.sizeof(L.BnPart), L.BnPart.dbeta.offset,
            ctypes.sizeof(L.PackItem), L.PackItem.slice_stride.offset,
            G.stats.offset, J.partials.offset, J.partial_rows.offset]
    from drn_b200 import parallel as PAR
    mine += [ctypes.sizeof(PAR.P2PComm), PAR.P2PComm.buf.offset, PAR.P2PComm.flags.offset, PAR.P2P_MAX_RANKS, PAR.P2P_FLAG_WORDS]
    assert [int(x) for x in out] == mine


def test_model_state_dict_contract_and_no_cpu_fallback():
    from model.main_model import mainModel
    m = mainModel(1301, S.config_namespace(stage=1))
    sd = m.state_dict()
    sp = spec_mod.state_dict_spec(S.default_config(stage=1))
    assert [k for k, _ in sp] == list(sd.keys())
    assert all(tuple(sd[k].shape) == tuple(s) for k, s in sp)
    m.load_state_dict(S.synth_state_dict(sp))
    assert sum(p.numel() for p in m.parameters()) == 45511378
    # attributes the reference's main.py touches (main.py:94, 126-133)
    assert m.query_encoder.embedding.weight.shape == (1302, 300)
    assert len(list(m.fcos.head.iou_scores.parameters())) == 6 and len(list(m.fcos.head.mix_fc.parameters())) == 4
    b = S.synth_batch(2, 32)
    with pytest.raises(RuntimeError, match="no CPU"):
        m(b["query_tokens"], b["query_length"], b["props_features"], b["props_start_end"], b["gt_start_end"], None, None)
    with pytest.raises(RuntimeError):
        m.backbone_net(torch.zeros(1))


def _emulate_rows(x, w, taps, par, t_out):
    """drn_gemm ROWS semantics in torch: x [B,T_in,C] viewed [B,T_in/par,par,C]; w [tap][N][K]."""
    B, Tin, Cn = x.shape
    xv = x.view(B, Tin // par, par, Cn)
    out = torch.zeros(B, t_out, w.shape[1])
    for sh, p, wt in taps:
        for t in range(t_out):
            ts = t + sh
            if 0 <= ts < Tin // par:
                out[:, t] += xv[:, ts, p] @ w[wt].t()
    return out


def test_tap_tables_are_the_reference_convolutions():
    torch.manual_seed(0)
    B, T, Ci, Co = 2, 16, 5, 7
    x = torch.randn(B, T, Ci)
    w = torch.randn(Co, Ci, 3)
    wt = w.permute(2, 0, 1).contiguous()  # tap-major [k][O][C] as drn_pack_conv_weight stores it
    for stride, taps in ((1, dense.K3), (2, dense.K3S2)):
        ref = F.conv1d(x.permute(0, 2, 1), w, stride=stride, padding=1).permute(0, 2, 1)
        got = _emulate_rows(x, wt, taps, stride, T // stride)
        assert torch.allclose(got, ref, atol=1e-5)
    # data gradients: stride 1, and stride 2 through the two output parities
    dy = torch.randn(B, T, Co)
    xx = x.clone().requires_grad_(True)
    F.conv1d(xx.permute(0, 2, 1), w, stride=1, padding=1).permute(0, 2, 1).mul(dy).sum().backward()
    wdg = w.permute(2, 1, 0).contiguous()  # [k][C][O]: "N" = C_in rows, "K" = C_out
    got = _emulate_rows(dy, wdg, dense.K3_DGRAD, 1, T)
    assert torch.allclose(got, xx.grad, atol=1e-5)
    dy2 = torch.randn(B, T // 2, Co)
    xx = x.clone().requires_grad_(True)
    F.conv1d(xx.permute(0, 2, 1), w, stride=2, padding=1).permute(0, 2, 1).mul(dy2).sum().backward()
    got = torch.zeros(B, T, Ci)
    for par in (0, 1):
        got[:, par::2] = _emulate_rows(dy2, wdg, dense.S2_DGRAD[par], 1, T // 2)
    assert torch.allclose(got, xx.grad, atol=1e-5)


def test_postprocess_list_assembly():
    """Host half of the eval post-processing (reference inference.py:167-215) on the fixed-shape arrays drn_postprocess writes:
    level concatenation in order, per-level `level` lists, and the fallback detection when nothing passed."""
    from model.inference import assemble
    B, nl, K = 3, 3, 4
    det = torch.arange(B * nl * K * 2, dtype=torch.float32).view(B, nl, K, 2) / 100
    score = torch.arange(B * nl * K, dtype=torch.float32).view(B, nl, K) / 50
    loc = score + 0.25
    count = torch.tensor([[2, 0, 1], [0, 0, 0], [4, 4, 4]], dtype=torch.int32)
    out = assemble(det, score, loc, count)
    assert out[0]["detections"].tolist() == torch.cat([det[0, 0, :2], det[0, 2, :1]]).tolist()
    assert out[0]["scores"].tolist() == torch.cat([score[0, 0, :2], score[0, 2, :1]]).tolist()
    assert out[0]["level"] == [[0, 0], [], [2]] and out[0]["labels"] == []
    assert out[1]["detections"].tolist() == [[0.0, 1.0]] and out[1]["level"] == [[-1]]
    assert out[1]["scores"].tolist() == [1.0] and out[1]["locations"].tolist() == [0.5]
    assert out[2]["detections"].shape == (12, 2) and out[2]["locations"].tolist() == loc[2].reshape(-1).tolist()


def test_reference_checkpoint_round_trip():
    """main.py:104-111 / 369-373: `module.`-prefixed DataParallel state_dicts load into the bare model (partial, key-matched) and
    checkpoints written for the reference carry the prefix and every reference key."""
    from drn_b200.checkpoint import load_reference_checkpoint, reference_checkpoint
    from model.main_model import mainModel
    sp = spec_mod.state_dict_spec(S.default_config(stage=1))
    src = S.synth_state_dict(sp)
    ck = {"epoch": 3, "state_dict": {"module." + k: v for k, v in src.items()}, "loss": 0.1, "top1": 44.7, "top5": 87.9}
    ck["state_dict"]["module.some_removed_layer.weight"] = torch.zeros(3)  # key-matched partial load skips unknown keys
    m = mainModel(1301, S.config_namespace(stage=1))
    loaded, skipped = load_reference_checkpoint(m, ck)
    assert skipped == ["module.some_removed_layer.weight"] and len(loaded) == len(src)
    got = m.state_dict()
    assert all(torch.equal(got[k], v) for k, v in src.items())
    out = reference_checkpoint(m, epoch=4)
    assert sorted(out["state_dict"]) == sorted("module." + k for k in src) and out["epoch"] == 4
    bad = {"state_dict": {"module.prop_fc.bias": torch.zeros(7)}}
    with pytest.raises(RuntimeError, match="shape"):
        load_reference_checkpoint(m, bad)


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the reference algorithm on the host cores: the one place besides tests / smoke where the
    oracle may run) prints ONE JSON line with the own arm's metric / unit / workload and the keys the driver reads."""
    import json
    import sys
    env = dict(os.environ, OMP_NUM_THREADS=str(os.cpu_count()))
    out = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, check=True, env=env, cwd=REPO, timeout=600).stdout.strip().splitlines()
    assert len(out) == 1
    d = json.loads(out[0])
    import importlib.util
    spec = importlib.util.spec_from_file_location("_bench", os.path.join(REPO, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    assert d["impl"] == "reference" and d["metric"] == b.METRIC and d["unit"] == "pairs/s" and d["higher_is_better"] is True
    assert d["config"] == b.workload_config(1) and d["value"] > 0 and d["vs_baseline"] is None  # the own arm emits the same dict
    assert d["steps_requested"] == 1 and d["steps"] == 1 and isinstance(d["steps_capped"], bool)
    # kind "reference" when build() has packed oracle/_ref/drn_reference.zip (the unmodified reference runs), else the port
    packed = os.path.isfile(os.path.join(REPO, "oracle", "_ref", "drn_reference.zip"))
    assert d["cpu_baseline"]["kind"] == ("reference" if packed else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_gradients_are_attached_as_views_and_still_accumulate(monkeypatch):
    """_DenseFn.backward hands the path's gradients to the parameters as views of the flat buffer (no autograd clone).  With
    CPU stand-ins for the two device schedules: the view is attached, zero_grad/None + a new backward overwrites it, and a
    backward WITHOUT zeroing in between accumulates old + new (the buffer is overwritten by every backward, so the old
    gradient must have been detached from it first)."""
    import types
    from model import main_model as MM

    w = torch.nn.Parameter(torch.arange(6.0).view(2, 3))
    v = torch.nn.Parameter(torch.ones(4))
    model = types.SimpleNamespace(_trainable_names=["w", "v"], use_graphs=False, _dp=None, _tensor_dict=lambda: {"w": w, "v": v})
    flat = torch.zeros(16)
    grads = {"w": flat[0:6].view(2, 3), "v": flat[8:12]}
    path = types.SimpleNamespace(losses=torch.zeros(8), poison=torch.zeros(1), flat_storages={flat.untyped_storage().data_ptr()})
    scale = {"k": 1.0}

    def fake_forward(path_, p, training, use_graphs, *inputs):
        path_.losses[:3] = torch.tensor([1.0, 2.0, 3.0])

    def fake_backward(path_, p, names, upstream, use_graphs, dp=None, flat_cache=None):
        flat.zero_()                       # what every real backward does: the buffer is rewritten
        grads["w"].copy_(scale["k"] * upstream[0] * torch.ones(2, 3))
        grads["v"].copy_(scale["k"] * upstream[1] * torch.full((4,), 2.0))
        return flat, grads
    monkeypatch.setattr(MM, "_run_forward", fake_forward)
    monkeypatch.setattr(MM, "_run_backward", fake_backward)
    dummy = torch.zeros(1)

    def step():
        losses = MM._DenseFn.apply(model, path, True, dummy, dummy, dummy, dummy, dummy, w, v)
        (losses[0] + 3.0 * losses[1]).backward()
    step()
    assert w.grad.untyped_storage().data_ptr() == flat.untyped_storage().data_ptr()   # a view, not a clone
    assert torch.equal(w.grad, torch.ones(2, 3)) and torch.equal(v.grad, torch.full((4,), 6.0))
    w.grad = v.grad = None
    scale["k"] = 2.0
    step()
    assert torch.equal(w.grad, torch.full((2, 3), 2.0)) and torch.equal(v.grad, torch.full((4,), 12.0))
    scale["k"] = 5.0
    step()                                  # no zeroing in between: 2 + 5, 12 + 30
    assert torch.equal(w.grad, torch.full((2, 3), 7.0)) and torch.equal(v.grad, torch.full((4,), 42.0))
    assert w.grad.untyped_storage().data_ptr() != flat.untyped_storage().data_ptr()
    # DRN_GRAD_VIEWS=0: autograd receives the gradients and clones them
    monkeypatch.setenv("DRN_GRAD_VIEWS", "0")
    w.grad = v.grad = None
    scale["k"] = 1.0
    step()
    assert torch.equal(w.grad, torch.ones(2, 3)) and w.grad.untyped_storage().data_ptr() != flat.untyped_storage().data_ptr()


def test_backward_after_another_forward_of_the_same_shape_raises(monkeypatch):
    """The backward reads the activations of its forward from the path's static buffers: a second forward of the same shape
    in between (eval pass, two losses) must raise instead of silently differentiating the wrong batch."""
    import types
    from model import main_model as MM

    w = torch.nn.Parameter(torch.ones(3))
    model = types.SimpleNamespace(_trainable_names=["w"], use_graphs=False, _dp=None, _tensor_dict=lambda: {"w": w})
    flat = torch.zeros(8)
    path = types.SimpleNamespace(losses=torch.zeros(8), poison=torch.zeros(1), flat_storages=set())
    monkeypatch.setattr(MM, "_run_forward", lambda path_, p, *a: path_.losses.__setitem__(slice(0, 3), torch.ones(3)))
    monkeypatch.setattr(MM, "_run_backward", lambda path_, p, names, up, ug, dp=None, flat_cache=None: (flat, {"w": flat[:3]}))
    dummy = torch.zeros(1)
    l1 = MM._DenseFn.apply(model, path, True, dummy, dummy, dummy, dummy, dummy, w)
    l2 = MM._DenseFn.apply(model, path, True, dummy, dummy, dummy, dummy, dummy, w)
    l2.sum().backward()  # the latest forward: fine
    with pytest.raises(RuntimeError, match="another forward"):
        l1.sum().backward()


# ---- launch planner of the persistent contraction kernel (drn_gemm_schedule_probe: host arithmetic only) ----------------------
def _probe(tiles, nk, pairs=74, ws=1, mode=1):
    import ctypes as C
    from drn_b200 import lib as L
    lib = L.load()
    n = len(tiles)
    kind, quota, static_tiles = C.c_int(), C.c_int(), C.c_int()
    counts, lists = (C.c_ubyte * 80)(), (C.c_ushort * (80 * 16))()
    cl = lib.drn_gemm_schedule_probe(n, (C.c_int * n)(*tiles), (C.c_int * n)(*nk), pairs, ws, mode, C.byref(kind), C.byref(quota),
                                     C.byref(static_tiles), counts, lists)
    per_pair = [list(lists)[c * 16:c * 16 + counts[c]] for c in range(max(cl, 0))]
    return cl, kind.value, quota.value, static_tiles.value, per_pair


def test_planner_balanced_lists_cover_every_tile_once_and_beat_round_robin():
    """Groups whose problems have tiles of different lengths get host-balanced (longest-first) tile lists: every tile exactly
    once, no pair above 16 tiles, makespan never above plain round-robin -- on the shapes of the DRN backward launches."""
    cases = {"conv0 bwd": ([204, 544], [32, 12]), "FPN layer bwd": ([84, 112], [32, 24]), "FPN inner bwd": ([48, 96], [16, 8]),
             "towers bwd": ([168, 112], [48, 32]), "FPN inner fwd": ([16, 32, 64], [16, 8, 4]), "conv1 bwd": ([24, 16, 16], [16, 16, 8])}
    for name, (tiles, nk) in cases.items():
        order = sorted(range(len(tiles)), key=lambda k: -nk[k])  # the launcher sorts problems by decreasing tile length
        tiles, nk = [tiles[k] for k in order], [nk[k] for k in order]
        cl, kind, quota, static_tiles, lists = _probe(tiles, nk)
        total = sum(tiles)
        assert cl == min(74, total)
        if total <= cl:
            continue
        assert kind == 1, name
        seen = sorted(t for lst in lists for t in lst)
        assert seen == list(range(total)), name
        assert max(len(lst) for lst in lists) <= 16
        starts = [sum(tiles[:k]) for k in range(len(tiles) + 1)]
        cost = lambda t: nk[max(k for k in range(len(tiles)) if t >= starts[k])]  # noqa: E731
        balanced = max(sum(cost(t) for t in lst) for lst in lists)
        rr = max(sum(cost(t) for t in range(c, total, cl)) for c in range(cl))
        ideal = sum(n * w for n, w in zip(tiles, nk)) / cl
        assert balanced <= rr and balanced < ideal + max(nk), (name, balanced, rr, ideal)


def test_planner_hybrid_only_where_it_pays_and_static_without_workspace():
    # prop_fc weight gradient: 256 tiles of 128 iterations = 3.46 waves -> 3 waves static, the last 34 tiles in ranges of 59
    assert _probe([256], [128]) == (74, 2, 59, 222, [[]] * 74)
    # towers / FPN layer forward (16 / 11 iterations to gain: less than a fold costs), conv0 forward: static
    assert _probe([224], [24])[:4] == (74, 0, 0, 0)
    assert _probe([112], [24])[:4] == (74, 0, 0, 0)
    assert _probe([64], [102])[:4] == (64, 0, 0, 0)
    # no workspace (data-parallel weight-gradient chunks: 64 tiles must stay on 64 pairs) or DRN_SCHEDULE=static: round-robin
    assert _probe([64], [128], ws=0)[:4] == (64, 0, 0, 0)
    assert _probe([256], [128], ws=0)[:4] == (74, 0, 0, 0)
    assert _probe([256], [128], mode=0)[:4] == (74, 0, 0, 0)
    # full stream-K: every pair gets ceil(total / pairs) iterations
    cl, kind, quota, static_tiles, _ = _probe([256], [128], mode=2)
    assert (cl, kind, static_tiles) == (74, 3, 0) and quota == -(-256 * 128 // 74)
    # uniform tiles in several problems: round-robin is already balanced -> no lists
    assert _probe([64, 32, 16], [24, 24, 24], ws=0)[:2] == (74, 0)


def test_weight_gradient_k_splits_never_exceed_the_k_blocks():
    """DensePath._wgrad_split: ~2048 contraction rows per tile, twice as fine for conv1 / the FPN laterals, and never more slices
    than 64-row K-blocks of the contraction (drn_gemm rejects a slice that would stay unwritten: gemm.cu prepare())."""
    import types
    from drn_b200.dense import DensePath

    def kblocks(b, t):  # gemm.cu: Rk = 64 or pow2_ceil(T) < 64, Bbk samples per block
        if t >= 64:
            return b * ((t + 63) // 64)
        rk = 1
        while rk < t:
            rk *= 2
        return (b + 64 // rk - 1) // (64 // rk)

    for b in (1, 3, 4, 32, 256):
        for t in (8, 16, 32, 64, 128, 256, 512):
            for prefix in ("backbone_net.forward_conv0", "backbone_net.forward_conv1", "fpn.fpn_inner2", "fpn.fpn_layer1", "fcos.head.towers"):
                blk = types.SimpleNamespace(prefix=prefix, rows=b * t, t_out=t)
                s = DensePath._wgrad_split(blk)
                assert 1 <= s <= 16 and s <= max(1, kblocks(b, t)), (b, t, prefix, s)
    full = lambda p: DensePath._wgrad_split(types.SimpleNamespace(prefix=p, rows=32 * 256, t_out=256))  # noqa: E731
    assert full("backbone_net.forward_conv0") == 4 and full("fpn.fpn_inner1") == 8
    assert DensePath._wgrad_split(types.SimpleNamespace(prefix="backbone_net.forward_conv1", rows=32 * 128, t_out=128)) == 4
