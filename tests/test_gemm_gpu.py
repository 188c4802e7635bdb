"""drn_gemm (tcgen05 split-BF16 contraction) against a torch fp64 reference of the same operands, and against the
library's fp32 CUDA-core checker engine.  Shapes cover every layer form of the path: linear, k3 conv stride 1 / 2,
data gradients (MN-major weights, parity-split outputs), weight gradients (MN-major both, split-K), ragged tiles."""
import pytest
import torch

from drn_b200 import lib as L
from drn_b200 import ops
from drn_b200.planes import Planes

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV)


def ref_rows(a, w, taps, stride_par, B, T, N, K, b_mn, a_c0=0):
    """a: Planes [B,Ta,C]; w: Planes [ntapsW, rows, cols].  Returns fp64 [B,T,N]."""
    A = a.to_float().double()
    W = w.to_float().double()
    P = stride_par
    Ta = A.shape[1] // P
    A = A.view(B, Ta, P, -1)
    out = torch.zeros(B, T, N, dtype=torch.float64, device=DEV)
    for (sh, par, wt) in taps:
        src = torch.zeros(B, T, K, dtype=torch.float64, device=DEV)
        lo, hi = max(0, -sh), min(T, Ta - sh)
        if hi > lo:
            src[:, lo:hi] = A[:, lo + sh:hi + sh, par, a_c0:a_c0 + K]
        Wt = W[wt, :K, :N] if b_mn else W[wt, :N, :K].t()
        out += src @ Wt
    return out


def run_rows(B, T, Cin, N, taps, P=1, b_mn=0, engine=0, nprod=3, seed=0, a_c0=0, K=None, dbg=(0, 0, 0)):
    K = K or Cin
    a = Planes.from_float(_rand(B, T * P, Cin, seed=seed))
    ntw = max(t[2] for t in taps) + 1
    w = Planes.from_float(_rand(ntw, K if b_mn else N, N if b_mn else K, seed=seed + 1, scale=K ** -0.5))
    out = torch.full((B, T, N), float("nan"), device=DEV)
    ops.gemm(L.GEMM_ROWS, a.desc(P), w.desc(), B, T, N, K=K, taps=taps, b_mn=b_mn, a_c0=a_c0, nprod=nprod, out=out,
             engine=engine, dbg=dbg)
    torch.cuda.synchronize()
    ref = ref_rows(a, w, taps, P, B, T, N, K, b_mn, a_c0)
    err = (out.double() - ref).abs().max().item() / ref.abs().max().item()
    return err, out, ref


K1 = ((0, 0, 0),)
K3 = ((-1, 0, 0), (0, 0, 1), (1, 0, 2))
K3S2 = ((-1, 1, 0), (0, 0, 1), (0, 1, 2))  # input viewed [T/2][2]: row 2t+r-1


@pytest.mark.parametrize("engine", [1, 3, 2])  # fp32 checker, one-tile-per-CTA tcgen05, persistent CTA-pair tcgen05
@pytest.mark.parametrize("case", [
    dict(B=2, T=256, Cin=128, N=128, taps=((0, 0, 0),)),
    dict(B=2, T=256, Cin=256, N=256, taps=K3),
    dict(B=3, T=128, Cin=192, N=512, taps=K3),
    dict(B=5, T=64, Cin=128, N=256, taps=K3),            # 2 samples per 128-row tile, ragged batch
    dict(B=4, T=32, Cin=64, N=128, taps=K3),
    dict(B=3, T=8, Cin=64, N=64, taps=K3),
    dict(B=2, T=96, Cin=64, N=1000, taps=K3),            # ragged T and N
    dict(B=2, T=128, Cin=128, N=256, taps=K3S2, P=2),    # stride-2 conv through the parity view
    dict(B=2, T=16, Cin=128, N=512, taps=K3S2, P=2),
    dict(B=2, T=256, Cin=256, N=128, taps=K3, b_mn=1),   # data gradient: weights [tap][K][N]
    dict(B=2, T=64, Cin=512, N=4352, taps=K3, b_mn=1),
    dict(B=2, T=256, Cin=320, N=128, taps=((0, 0, 0),), a_c0=64, K=256),
])
def test_rows(case, engine):
    err, _, _ = run_rows(engine=engine, **case)
    assert err < 2e-5, err


def test_rows_single_pass_is_bf16_accurate():
    err, _, _ = run_rows(2, 256, 256, 256, K3, nprod=1)
    assert 1e-5 < err < 2e-2, err


def run_wgrad(B, T, Co, Ci, taps, P=1, engine=0, split_k=1, seed=0, dbg=(0, 0, 0)):
    dy = Planes.from_float(_rand(B, T, Co, seed=seed))
    x = Planes.from_float(_rand(B, T * P, Ci, seed=seed + 1))
    ntap = len(taps)
    out = torch.zeros(ntap, Co, Ci, device=DEV)
    ops.gemm(L.GEMM_WGRAD, dy.desc(), x.desc(P), B, T, Ci, M=Co, taps=taps, out=out, out_ld=Ci,
             out_tap_stride=Co * Ci, out_mode=L.OUT_ATOMIC if split_k > 1 else L.OUT_STORE, split_k=split_k,
             engine=engine, dbg=dbg)
    torch.cuda.synchronize()
    DY = dy.to_float().double()
    X = x.to_float().double().view(B, T, P, Ci)
    ref = torch.zeros(ntap, Co, Ci, dtype=torch.float64, device=DEV)
    for (sh, par, wt) in taps:
        src = torch.zeros(B, T, Ci, dtype=torch.float64, device=DEV)
        lo, hi = max(0, -sh), min(T, T - sh)
        src[:, lo:hi] = X[:, lo + sh:hi + sh, par]
        ref[wt] = torch.einsum("bto,btc->oc", DY, src)
    err = (out.double() - ref).abs().max().item() / ref.abs().max().item()
    return err


@pytest.mark.parametrize("engine", [1, 3, 2])  # fp32 checker, one-tile-per-CTA tcgen05, persistent CTA-pair tcgen05
@pytest.mark.parametrize("case", [
    dict(B=2, T=256, Co=128, Ci=128, taps=((0, 0, 0),)),
    dict(B=2, T=128, Co=256, Ci=320, taps=K3),
    dict(B=5, T=32, Co=192, Ci=256, taps=K3),
    dict(B=3, T=8, Co=64, Ci=64, taps=K3),
    dict(B=2, T=64, Co=512, Ci=256, taps=K3S2, P=2),
    dict(B=4, T=96, Co=128, Ci=128, taps=K3),
])
def test_wgrad(case, engine):
    err = run_wgrad(engine=engine, **case)
    assert err < 2e-5, err


@pytest.mark.parametrize("engine", [3, 2])
def test_wgrad_split_k(engine):
    err = run_wgrad(8, 128, 256, 256, K3, split_k=4, engine=engine)
    assert err < 2e-5, err


@pytest.mark.parametrize("engine", [3, 2])
def test_many_tiles_persistent(engine):
    """More tiles than SM pairs: the persistent kernel walks several tiles per cluster (TMEM double buffering, pipeline
    state carried across tiles)."""
    err, _, _ = run_rows(16, 256, 128, 2048, K3, engine=engine)
    assert err < 2e-5, err
    err = run_wgrad(16, 256, 1024, 1536, K3, engine=engine)
    assert err < 2e-5, err


@pytest.mark.parametrize("engine", [3, 2])
def test_epilogue_options(engine):
    B, T, Cin, N = 2, 128, 128, 256
    a = Planes.from_float(_rand(B, T, Cin, seed=3))
    w = Planes.from_float(_rand(1, N, Cin, seed=4, scale=Cin ** -0.5))
    bias = _rand(N, seed=5)
    q = _rand(B, N, seed=6)
    wide = torch.zeros(B, T, N + 64, device=DEV)
    pre = torch.empty(B, T, N, device=DEV)
    pl = Planes.zeros(B, T, N + 64, DEV)
    ops.gemm(L.GEMM_ROWS, a.desc(), w.desc(), B, T, N, K=Cin, out=wide, out_ld=N + 64, out_col0=32, bias=bias, rowscale=q,
             out2=pre, outp=pl, outp_col0=64, engine=engine)
    ref_pre = ref_rows(a, w, ((0, 0, 0),), 1, B, T, N, Cin, 0) + bias.double()
    ref = ref_pre * q.double()[:, None, :]
    s = ref.abs().max().item()
    assert (pre.double() - ref_pre).abs().max().item() / s < 2e-5
    assert (wide[:, :, 32:32 + N].double() - ref).abs().max().item() / s < 2e-5
    assert wide[:, :, :32].abs().max().item() == 0 and wide[:, :, 32 + N:].abs().max().item() == 0
    assert (pl.to_float()[:, :, 64:].double() - ref).abs().max().item() / s < 3e-5
    # accumulate mode
    ops.gemm(L.GEMM_ROWS, a.desc(), w.desc(), B, T, N, K=Cin, out=wide, out_ld=N + 64, out_col0=32, bias=bias, rowscale=q,
             out_mode=L.OUT_ADD, engine=engine)
    assert (wide[:, :, 32:32 + N].double() - 2 * ref).abs().max().item() / s < 4e-5
    # parity-strided output rows (stride-2 data gradient writes every other time step)
    out = torch.zeros(B, 2 * T, N, device=DEV)
    ops.gemm(L.GEMM_ROWS, a.desc(), w.desc(), B, T, N, K=Cin, out=out, out_T=2 * T, out_t_mul=2, out_t_add=1, engine=engine)
    ref0 = ref_rows(a, w, ((0, 0, 0),), 1, B, T, N, Cin, 0)
    assert (out[:, 1::2].double() - ref0).abs().max().item() / s < 2e-5
    assert out[:, 0::2].abs().max().item() == 0


def test_group_mixed_forms_and_split_slices():
    """drn_gemm_group: three ROWS problems of different sizes (the pyramid levels of a shared conv), a stride-2 data gradient
    pair and a weight gradient whose K-splits store their partial sums in slices (no atomics), all in ONE launch."""
    descs, checks, keep = [], [], []  # descriptors hold raw pointers: keep every operand alive until the launch
    w = Planes.from_float(_rand(3, 512, 256, seed=11, scale=(3 * 256) ** -0.5))
    for i, T in enumerate((256, 128, 64)):
        a = Planes.from_float(_rand(4, T, 256, seed=20 + i))
        keep.append(a)
        out = torch.full((4, T, 512), float("nan"), device=DEV)
        descs.append(ops.desc(L.GEMM_ROWS, a.desc(), w.desc(), 4, T, 512, K=256, taps=K3, out=out, engine=2))
        checks.append((out, ref_rows(a, w, K3, 1, 4, T, 512, 256, 0)))
    # weight gradient with 3 K-splits -> 3 slices, summed by the caller
    B, T, Co, Ci = 6, 128, 256, 320
    dy, x = Planes.from_float(_rand(B, T, Co, seed=31)), Planes.from_float(_rand(B, T, Ci, seed=32))
    ws = torch.full((3, 3, Co, Ci), float("nan"), device=DEV)
    descs.append(ops.desc(L.GEMM_WGRAD, dy.desc(), x.desc(), B, T, Ci, M=Co, taps=K3, out=ws[0], out_ld=Ci,
                          out_tap_stride=Co * Ci, out_split_stride=ws.stride(0), split_k=3, engine=2))
    DY, X = dy.to_float().double(), x.to_float().double()
    ref_w = torch.zeros(3, Co, Ci, dtype=torch.float64, device=DEV)
    for (sh, par, wt) in K3:
        src = torch.zeros(B, T, Ci, dtype=torch.float64, device=DEV)
        lo, hi = max(0, -sh), min(T, T - sh)
        src[:, lo:hi] = X[:, lo + sh:hi + sh]
        ref_w[wt] = torch.einsum("bto,btc->oc", DY, src)
    assert ops.gemm_group(descs) == 1
    torch.cuda.synchronize()
    for out, ref in checks:
        assert (out.double() - ref).abs().max().item() / ref.abs().max().item() < 2e-5
    assert not torch.isnan(ws).any()
    assert (ws.double().sum(0) - ref_w).abs().max().item() / ref_w.abs().max().item() < 2e-5
    # deterministic: a second launch reproduces every slice bit for bit
    ws2 = ws.clone()
    ops.gemm_group(descs)
    torch.cuda.synchronize()
    assert torch.equal(ws, ws2)


def test_group_rejects_bad_arguments():
    a = Planes.from_float(_rand(2, 128, 128, seed=1))
    w = Planes.from_float(_rand(1, 128, 128, seed=2))
    out = torch.zeros(2, 128, 128, device=DEV)
    d = ops.desc(L.GEMM_ROWS, a.desc(), w.desc(), 2, 128, 128, K=128, out=out, engine=2)
    arr = (L.GemmDesc * 7)(*([d] * 7))
    assert L.load().drn_gemm_group(7, arr, L.stream_ptr()) == -1
    d1 = ops.desc(L.GEMM_ROWS, a.desc(), w.desc(), 2, 128, 128, K=128, out=out, engine=1)
    arr = (L.GemmDesc * 1)(d1)
    assert L.load().drn_gemm_group(1, arr, L.stream_ptr()) == -1
    # more K-splits than K-blocks would leave slices unwritten
    dy, x = Planes.from_float(_rand(1, 64, 128, seed=3)), Planes.from_float(_rand(1, 64, 128, seed=4))
    ws = torch.zeros(4, 1, 128, 128, device=DEV)
    d2 = ops.desc(L.GEMM_WGRAD, dy.desc(), x.desc(), 1, 64, 128, M=128, out=ws[0], out_ld=128, out_tap_stride=128 * 128,
                  out_split_stride=ws.stride(0), split_k=4, engine=2)
    arr = (L.GemmDesc * 1)(d2)
    assert L.load().drn_gemm_group(1, arr, L.stream_ptr()) == -1


def test_rows_k_split_slices():
    """conv0-forward shape class: one N tile, long K (3 taps x 1024): the K-split of a ROWS problem stores two slices whose sum is
    the convolution."""
    B, T, Cin, N = 8, 128, 1024, 256
    a = Planes.from_float(_rand(B, T, Cin, seed=41))
    w = Planes.from_float(_rand(3, N, Cin, seed=42, scale=(3 * Cin) ** -0.5))
    ys = torch.full((2, B, T, N), float("nan"), device=DEV)
    d = ops.desc(L.GEMM_ROWS, a.desc(), w.desc(), B, T, N, K=Cin, taps=K3, out=ys[0], out_split_stride=ys.stride(0), split_k=2,
                 engine=2)
    assert ops.gemm_group([d]) == 1
    torch.cuda.synchronize()
    ref = ref_rows(a, w, K3, 1, B, T, N, Cin, 0)
    assert not torch.isnan(ys).any()
    assert float(ys[1].abs().max()) > 0
    assert (ys.double().sum(0) - ref).abs().max().item() / ref.abs().max().item() < 2e-5


@pytest.mark.parametrize("B,T,N", [(4, 256, 512), (3, 96, 256), (5, 32, 512), (2, 200, 256)])
def test_fused_batchnorm_partial_sums(B, T, N):
    """drn_gemm_t.stats: the epilogue's per-32-row partial column sums (after the bias, padding rows excluded) add up to the
    column sums / sums of squares of the conv output -- ragged last tiles (T = 96, 200), several samples per tile (T = 32) and an
    odd number of 128-row sub-tiles -- and drn_bn_stats_multi reduces them to the same BatchNorm coefficients and running
    statistics as its pass over y."""
    import ctypes as C
    Cin = 256
    a = Planes.from_float(_rand(B, T, Cin, seed=51))
    w = Planes.from_float(_rand(3, N, Cin, seed=52, scale=(3 * Cin) ** -0.5))
    bias = _rand(N, seed=53)
    out = torch.full((B, T, N), float("nan"), device=DEV)
    d = ops.desc(L.GEMM_ROWS, a.desc(), w.desc(), B, T, N, K=Cin, taps=K3, out=out, bias=bias, engine=2)
    lib = L.load()
    rows = lib.drn_gemm_stats_rows(C.byref(d))
    assert rows >= (B * T + 31) // 32
    stats = torch.full((rows, 2, N), float("nan"), device=DEV)
    d.stats = stats.data_ptr()
    assert ops.gemm_group([d]) == 1
    torch.cuda.synchronize()
    assert not torch.isnan(stats).any()
    y = out.double().view(-1, N)
    s = stats.double().sum(0)
    assert (s[0] - y.sum(0)).abs().max().item() <= 1e-5 * y.abs().sum(0).max().item()
    assert (s[1] - (y * y).sum(0)).abs().max().item() <= 1e-5 * (y * y).sum(0).max().item()
    # the statistics finaliser: partial sums vs a pass over y (same module state before each)
    def job(partials):
        j = L.BnJob()
        j.y, j.B, j.T, j.C, j.nparts = out.data_ptr(), B, T, N, 1
        st = dict(gamma=_rand(N, seed=54), beta=_rand(N, seed=55), rm=torch.zeros(N, device=DEV), rv=torch.ones(N, device=DEV),
                  nbt=torch.zeros(1, dtype=torch.int64, device=DEV), coef=torch.zeros(5, N, device=DEV),
                  sums=torch.zeros(2, N, dtype=torch.float64, device=DEV), cnt=torch.zeros(1, dtype=torch.int32, device=DEV))
        p = j.parts[0]
        p.c0, p.n, p.gamma, p.beta = 0, N, st["gamma"].data_ptr(), st["beta"].data_ptr()
        p.running_mean, p.running_var, p.num_batches_tracked = st["rm"].data_ptr(), st["rv"].data_ptr(), st["nbt"].data_ptr()
        j.coef, j.sums, j.counter = st["coef"].data_ptr(), st["sums"].data_ptr(), st["cnt"].data_ptr()
        if partials:
            j.partials, j.partial_rows = stats.data_ptr(), rows
        arr = (L.BnJob * 1)(j)
        L.check(lib.drn_bn_stats_multi(1, arr, C.c_float(0.1), C.c_float(1e-5), 1, L.stream_ptr()), "bn_stats")
        torch.cuda.synchronize()
        return st
    ref, got = job(False), job(True)
    for k in ("coef", "rm", "rv"):
        assert (got[k] - ref[k]).abs().max().item() <= 2e-6 * max(ref[k].abs().max().item(), 1.0), k
    assert int(got["nbt"]) == 1 and float(got["sums"].abs().max()) == 0.0 and int(got["cnt"]) == 0
    # wrong use is rejected: statistics with a K-split / through the one-tile-per-CTA engine
    d3 = ops.desc(L.GEMM_ROWS, a.desc(), w.desc(), B, T, N, K=Cin, taps=K3, out=out, engine=3)
    d3.stats = stats.data_ptr()
    assert lib.drn_gemm(C.byref(d3), L.stream_ptr()) == -1


def test_hybrid_schedule_splits_only_the_last_wave():
    """80 tiles of 48 k-iterations on 74 SM pairs: the first 74 run whole (bit-identical to the static schedule), the k-iterations
    of the last 6 are cut into ranges over all pairs and folded through the workspace -- against fp64, with bias and BatchNorm
    partial sums behind the fold, flags re-armed, bit-exact on a repeat; and a weight gradient with K-split slices (120 tiles)."""
    import ctypes as C
    assert ops.SCHEDULE == "hybrid"
    B, T, Cin, N = 20, 256, 1024, 1024
    a = Planes.from_float(_rand(B, T, Cin, seed=401))
    w = Planes.from_float(_rand(3, N, Cin, seed=402, scale=(3 * Cin) ** -0.5))
    bias = _rand(N, seed=403)
    y = torch.full((B, T, N), float("nan"), device=DEV)
    d = ops.desc(L.GEMM_ROWS, a.desc(), w.desc(), B, T, N, K=Cin, taps=K3, out=y, bias=bias, engine=2)
    rows = L.load().drn_gemm_stats_rows(C.byref(d))
    stats = torch.full((rows, 2, N), float("nan"), device=DEV)
    d.stats = stats.data_ptr()
    ops.gemm_group([d])
    assert _flags_clear()
    ref = ref_rows(a, w, K3, 1, B, T, N, Cin, 0) + bias.double()
    assert (y.double() - ref).abs().max().item() / ref.abs().max().item() < 2e-5
    yy = y.double().view(-1, N)
    st = stats.double().sum(0)
    assert (st[0] - yy.sum(0)).abs().max().item() <= 1e-5 * yy.abs().sum(0).max().item()
    assert (st[1] - (yy * yy).sum(0)).abs().max().item() <= 1e-5 * (yy * yy).sum(0).max().item()
    first, first_stats = y.clone(), stats.clone()
    ops.gemm_group([d])
    assert _flags_clear() and torch.equal(first, y) and torch.equal(first_stats, stats)
    ys = torch.full_like(y, float("nan"))
    d.out, d.stats = ys.data_ptr(), None
    _group_static([d])
    torch.cuda.synchronize()
    sm_pairs = L.load().drn_sm_count() // 2
    if sm_pairs == 74:  # tiles 0 .. 73 (row-major over 4 column tiles) never met the workspace
        whole = 74 // 4  # complete 256-row tile rows inside the first wave
        assert torch.equal(ys.view(-1, N)[:whole * 256], first.view(-1, N)[:whole * 256])
    assert (ys - first).abs().max().item() / ref.abs().max().item() < 2e-5
    # weight gradient: 4 x 5 tiles x 3 taps x 2 slices = 120 tiles of 64 iterations
    B, T, Co, Ci = 32, 256, 1024, 1280
    dy, x = Planes.from_float(_rand(B, T, Co, seed=411)), Planes.from_float(_rand(B, T, Ci, seed=412))
    ws = torch.full((2, 3, Co, Ci), float("nan"), device=DEV)
    dw = ops.desc(L.GEMM_WGRAD, dy.desc(), x.desc(), B, T, Ci, M=Co, taps=K3, out=ws[0], out_ld=Ci,
                  out_tap_stride=Co * Ci, out_split_stride=ws.stride(0), split_k=2, engine=2)
    ops.gemm_group([dw])
    assert _flags_clear() and not torch.isnan(ws).any()
    DY, X = dy.to_float().double(), x.to_float().double()
    for (sh, par, wt) in K3:
        src = torch.zeros(B, T, Ci, dtype=torch.float64, device=DEV)
        lo, hi = max(0, -sh), min(T, T - sh)
        src[:, lo:hi] = X[:, lo + sh:hi + sh]
        ref_w = torch.einsum("bto,btc->oc", DY, src)
        assert (ws[:, wt].double().sum(0) - ref_w).abs().max().item() / ref_w.abs().max().item() < 2e-5


def test_group_balanced_schedule_is_bit_identical_to_single_launches():
    """A group whose problems have tiles of different lengths and more tiles than SM pairs is walked on the host-balanced
    (longest-first) schedule; which pair computes a tile must not change a bit of it: every problem equals its own launch."""
    descs, outs, keep = [], [], []
    # weight gradient with 32-iteration tiles (2 x 2 tiles x 3 taps x 4 slices = 48 tiles)
    B, T, Co, Ci = 32, 256, 512, 512
    dy, x = Planes.from_float(_rand(B, T, Co, seed=301)), Planes.from_float(_rand(B, T, Ci, seed=302))
    ws = torch.full((4, 3, Co, Ci), float("nan"), device=DEV)
    descs.append(ops.desc(L.GEMM_WGRAD, dy.desc(), x.desc(), B, T, Ci, M=Co, taps=K3, out=ws[0], out_ld=Ci,
                          out_tap_stride=Co * Ci, out_split_stride=ws.stride(0), split_k=4, engine=2))
    outs.append(ws)
    # 1 x 1 convs with 4- and 16-iteration tiles on ragged rows (64 + 36 tiles), a 3-tap conv with 24-iteration tiles (32 tiles)
    for i, (b, t, cin, n, taps) in enumerate(((32, 250, 256, 512, K1), (17, 130, 1024, 1000, K1), (16, 256, 512, 512, K3))):
        a = Planes.from_float(_rand(b, t, cin, seed=310 + i))
        w = Planes.from_float(_rand(len(taps), n, cin, seed=320 + i, scale=(len(taps) * cin) ** -0.5))
        keep += [a, w]
        o = torch.full((b, t, n), float("nan"), device=DEV)
        descs.append(ops.desc(L.GEMM_ROWS, a.desc(), w.desc(), b, t, n, K=cin, taps=taps, out=o, engine=2))
        outs.append(o)
    assert ops.gemm_group(descs) == 1
    torch.cuda.synchronize()
    grouped = [o.clone() for o in outs]
    for o in outs:
        o.fill_(float("nan"))
    for d in descs:  # one problem per launch: plain round-robin
        arr = (L.GemmDesc * 1)(d)
        L.check(L.load().drn_gemm_group(1, arr, L.stream_ptr()), "drn_gemm_group (single)")
    torch.cuda.synchronize()
    for g, o in zip(grouped, outs):
        assert not torch.isnan(g).any() and torch.equal(g, o)


# ---- stream-K schedule (drn_gemm_group_ws; opt-in, DRN_STREAMK=1 in the product path): many partial tiles per owner, every
# ---- epilogue option behind a fold, static-schedule agreement, determinism, flags re-armed
def _group_static(descs):
    arr = (L.GemmDesc * len(descs))(*descs)
    L.check(L.load().drn_gemm_group(len(descs), arr, L.stream_ptr()), "drn_gemm_group (static schedule)")


def _flags_clear():
    torch.cuda.synchronize()
    return int(ops.workspace()[:8192].max()) == 0


def test_streamk_long_tiles_many_contributors():
    """Two 256 x 256 tiles of 96 k-iterations each (3 taps x 2048 channels): stream-K cuts them into ~48 ranges of 4 iterations,
    so each owner folds ~23 partial tiles -- against fp64, against the static schedule, bit-exact on a repeat, flags re-armed."""
    B, T, Cin, N = 2, 256, 2048, 256
    a = Planes.from_float(_rand(B, T, Cin, seed=61))
    w = Planes.from_float(_rand(3, N, Cin, seed=62, scale=(3 * Cin) ** -0.5))
    out = torch.full((B, T, N), float("nan"), device=DEV)
    d = ops.desc(L.GEMM_ROWS, a.desc(), w.desc(), B, T, N, K=Cin, taps=K3, out=out, engine=2)
    ops.gemm_group([d], streamk=True)
    assert _flags_clear()
    ref = ref_rows(a, w, K3, 1, B, T, N, Cin, 0)
    assert (out.double() - ref).abs().max().item() / ref.abs().max().item() < 2e-5
    first = out.clone()
    ops.gemm_group([d], streamk=True)
    assert _flags_clear() and torch.equal(first, out)
    stat = torch.full_like(out, float("nan"))
    d.out = stat.data_ptr()
    _group_static([d])
    torch.cuda.synchronize()
    # one accumulator over all 96 iterations (1152 dependent tensor-core accumulations): measurably less accurate than the sum
    # of short partial tiles folded in fp32 round-to-nearest
    assert (stat.double() - ref).abs().max().item() / ref.abs().max().item() < 6e-5
    assert (stat - first).abs().max().item() / ref.abs().max().item() < 6e-5


def test_streamk_ragged_remainder_mixed_group():
    """The shape class of the head backward: three pyramid levels of data gradients (K = 3 x 512 -> 24 iterations a tile) and
    weight gradients with K-split slices, tile counts that do not divide by the 74 SM pairs, ragged N and ragged rows."""
    descs, checks, keep = [], [], []
    w = Planes.from_float(_rand(3, 320, 512, seed=71, scale=(3 * 320) ** -0.5))       # [tap][K][N] for the data gradient
    wf = Planes.from_float(_rand(3, 1000, 320, seed=72, scale=(3 * 320) ** -0.5))     # [tap][N][K], ragged N = 1000
    for i, (B, T) in enumerate(((9, 200), (7, 100), (5, 52))):
        dy = Planes.from_float(_rand(B, T, 320, seed=80 + i))
        keep.append(dy)
        o1 = torch.full((B, T, 512), float("nan"), device=DEV)
        descs.append(ops.desc(L.GEMM_ROWS, dy.desc(), w.desc(), B, T, 512, K=320, taps=K3, b_mn=1, out=o1, engine=2))
        checks.append((o1, ref_rows(dy, w, K3, 1, B, T, 512, 320, 1)))
        if i < 2:
            o2 = torch.full((B, T, 1000), float("nan"), device=DEV)
            descs.append(ops.desc(L.GEMM_ROWS, dy.desc(), wf.desc(), B, T, 1000, K=320, taps=K3, out=o2, engine=2))
            checks.append((o2, ref_rows(dy, wf, K3, 1, B, T, 1000, 320, 0)))
    B, T, Co, Ci = 10, 128, 320, 448
    dy, x = Planes.from_float(_rand(B, T, Co, seed=91)), Planes.from_float(_rand(B, T, Ci, seed=92))
    ws = torch.full((4, 3, Co, Ci), float("nan"), device=DEV)
    descs.append(ops.desc(L.GEMM_WGRAD, dy.desc(), x.desc(), B, T, Ci, M=Co, taps=K3, out=ws[0], out_ld=Ci,
                          out_tap_stride=Co * Ci, out_split_stride=ws.stride(0), split_k=4, engine=2))
    DY, X = dy.to_float().double(), x.to_float().double()
    ref_w = torch.zeros(3, Co, Ci, dtype=torch.float64, device=DEV)
    for (sh, par, wt) in K3:
        src = torch.zeros(B, T, Ci, dtype=torch.float64, device=DEV)
        lo, hi = max(0, -sh), min(T, T - sh)
        src[:, lo:hi] = X[:, lo + sh:hi + sh]
        ref_w[wt] = torch.einsum("bto,btc->oc", DY, src)
    assert len(descs) == 6 and ops.gemm_group(descs, streamk=True) == 1
    assert _flags_clear()
    for out, ref in checks:
        assert not torch.isnan(out).any()
        assert (out.double() - ref).abs().max().item() / ref.abs().max().item() < 2e-5
    assert not torch.isnan(ws).any()
    assert (ws.double().sum(0) - ref_w).abs().max().item() / ref_w.abs().max().item() < 2e-5
    snap = [o.clone() for o, _ in checks] + [ws.clone()]
    ops.gemm_group(descs, streamk=True)
    assert _flags_clear()
    for s, t in zip(snap, [o for o, _ in checks] + [ws]):
        assert torch.equal(s, t)


def test_streamk_epilogue_options_behind_a_fold():
    """bias, second output, row gate, accumulate mode, plane output and BatchNorm partial sums all run AFTER the fold: 6 tiles of
    48 iterations on 74 pairs -> every tile is folded from several partial tiles."""
    import ctypes as C
    B, T, Cin, N = 3, 256, 1024, 512
    a = Planes.from_float(_rand(B, T, Cin, seed=101))
    w = Planes.from_float(_rand(3, N, Cin, seed=102, scale=(3 * Cin) ** -0.5))
    bias, q = _rand(N, seed=103), _rand(B, N, seed=104)
    base = _rand(B, T, N, seed=105)
    out, pre, pl = base.clone(), torch.empty(B, T, N, device=DEV), Planes.zeros(B, T, N, DEV)
    d = ops.desc(L.GEMM_ROWS, a.desc(), w.desc(), B, T, N, K=Cin, taps=K3, out=out, out_mode=L.OUT_ADD, bias=bias, rowscale=q,
                 out2=pre, outp=pl, engine=2)
    ops.gemm_group([d], streamk=True)
    assert _flags_clear()
    ref_pre = ref_rows(a, w, K3, 1, B, T, N, Cin, 0) + bias.double()
    ref = ref_pre * q.double()[:, None, :]
    s = ref.abs().max().item()
    assert (pre.double() - ref_pre).abs().max().item() / s < 2e-5
    assert (out.double() - base.double() - ref).abs().max().item() / s < 4e-5
    assert (pl.to_float().double() - ref).abs().max().item() / s < 3e-5
    y = torch.full((B, T, N), float("nan"), device=DEV)
    d2 = ops.desc(L.GEMM_ROWS, a.desc(), w.desc(), B, T, N, K=Cin, taps=K3, out=y, bias=bias, engine=2)
    rows = L.load().drn_gemm_stats_rows(C.byref(d2))
    stats = torch.full((rows, 2, N), float("nan"), device=DEV)
    d2.stats = stats.data_ptr()
    ops.gemm_group([d2], streamk=True)
    assert _flags_clear() and not torch.isnan(stats).any()
    yy = y.double().view(-1, N)
    st = stats.double().sum(0)
    assert (st[0] - yy.sum(0)).abs().max().item() <= 1e-5 * yy.abs().sum(0).max().item()
    assert (st[1] - (yy * yy).sum(0)).abs().max().item() <= 1e-5 * (yy * yy).sum(0).max().item()


def test_streamk_matches_static_on_full_size_layers():
    """prop_fc-forward and prop_fc-wgrad shape classes at a size with several tiles per SM pair: stream-K and the static
    schedule agree to fp32 summation-order noise."""
    old = ops.STREAMK
    ops.STREAMK = True
    try:
        err, out, ref = run_rows(8, 256, 1024, 2048, ((0, 0, 0),), engine=2)
        assert err < 2e-5, err
        err = run_wgrad(16, 256, 1024, 1536, ((0, 0, 0),), engine=2)
        assert err < 2e-5, err
    finally:
        ops.STREAMK = old
    assert _flags_clear()
