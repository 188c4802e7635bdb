"""Query-encoder kernels (drn_qe_forward / drn_qe_backward, drn_sgemm) through the C ABI against the CPU oracle
(oracle.drn_oracle.query_encoder = reference model/language_module.py:27-62) and torch autograd of that oracle.

Floating point, exact-fp32 FMA arithmetic on both sides: tolerance 1e-4 relative (max-norm) forward, 1e-3 rel-L2 on
gradients (sums of up to B*L*4H products re-associated differently).  Edge cases: ragged and UNSORTED lengths, length 1,
length = L, more token columns than the longest query, B > 32 (sample chunking), padding tokens inside the gradient."""
import ctypes as C

import pytest
import torch

from drn_b200 import lib as L
from drn_b200 import spec as spec_mod
from drn_b200 import synthetic as S
from drn_b200.dense import DensePath
from oracle import drn_oracle as O

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-30))


def _rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / max(float(b.norm()), 1e-30))


@pytest.fixture(scope="module")
def sd():
    return S.synth_state_dict(spec_mod.state_dict_spec(S.default_config(stage=1)))


def _tokens(B, Lcols, lengths, seed=0):
    g = torch.Generator().manual_seed(seed)
    tok = torch.zeros(B, Lcols, dtype=torch.int64)
    for b, n in enumerate(lengths):
        tok[b, :n] = torch.randint(1, 1302, (n,), generator=g)
    return tok


CASES = [
    ("sorted_b4", 4, 10, [10, 7, 3, 2]),
    ("unsorted_len1_b5", 5, 8, [3, 8, 1, 5, 1]),
    ("extra_columns_b3", 3, 12, [6, 4, 2]),
    ("chunked_b40", 40, 10, [1 + (7 * i) % 10 for i in range(40)]),
]


@pytest.mark.parametrize("name,B,Lc,lengths", CASES)
def test_query_encoder_forward_backward(sd, name, B, Lc, lengths):
    dev = torch.device("cuda")
    cfg = S.default_config(stage=1)
    tok, lens = _tokens(B, Lc, lengths), torch.tensor(lengths, dtype=torch.int64)
    leaf = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in sd.items() if k.startswith("query_encoder.")}
    cmds, Hs = O.query_encoder(leaf, tok, lens)
    # hidden units of relu(qInput(q_vector)) that sit within 1e-4 of the ReLU kink for some sample: a 1e-6 rounding
    # difference flips their mask, so their rows of d qInput.{weight,bias} are excluded from the comparison
    with torch.no_grad():
        vq = torch.cat([Hs[:, 0], Hs[torch.arange(B), lens - 1]], dim=-1)
        pre = torch.nn.functional.linear(vq, leaf["query_encoder.qInput.weight"], leaf["query_encoder.qInput.bias"])
        safe_units = ~(pre.abs() < 1e-4).any(dim=0)
    assert int(safe_units.sum()) > 400
    g = torch.Generator().manual_seed(1)
    dcmd = [torch.randn(B, 1024, generator=g) for _ in range(3)]
    sum((c * d).sum() for c, d in zip(cmds, dcmd)).backward()

    path = DensePath(cfg, B, 32, dev, L=Lc)
    p = {k: v.to(dev) for k, v in sd.items()}
    path.tokens.copy_(tok)
    path.lengths.copy_(lens)
    L.check(L.load().drn_qe_forward(C.byref(path._qe_desc(p)), L.stream_ptr()), "qe_forward")
    for t in range(3):
        assert _rel(path.cmd[t], cmds[t].detach()) <= 1e-4, (name, t, _rel(path.cmd[t], cmds[t].detach()))
    grads = {k: torch.zeros_like(p[k]) for k in leaf if "textualAttention" not in k}
    for t in range(3):
        path.dcmd[t].copy_(dcmd[t])
    L.check(L.load().drn_qe_backward(C.byref(path._qe_desc(p, grads)), L.stream_ptr()), "qe_backward")
    torch.cuda.synchronize()
    for k, gk in grads.items():
        ref = leaf[k].grad
        assert ref is not None, k
        if float(ref.norm()) < 1e-5:  # softmax is shift invariant: d cmd_inter2logits.bias == 0
            assert float(gk.norm()) < 1e-4, k
            continue
        if k in ("query_encoder.qInput.weight", "query_encoder.qInput.bias"):
            gk, ref = gk.cpu()[safe_units], ref[safe_units]
        assert _rel_l2(gk, ref) <= 1e-3, (name, k, _rel_l2(gk, ref))
    assert float(grads["query_encoder.embedding.weight"][0].abs().max()) == 0.0  # padding_idx row
    # second call on the same workspace gives identical results (self-resetting counters / carried state)
    first = [c.clone() for c in path.cmd]
    L.check(L.load().drn_qe_forward(C.byref(path._qe_desc(p)), L.stream_ptr()), "qe_forward")
    g2 = {k: torch.zeros_like(v) for k, v in grads.items()}
    L.check(L.load().drn_qe_backward(C.byref(path._qe_desc(p, g2)), L.stream_ptr()), "qe_backward")
    torch.cuda.synchronize()
    assert all(torch.equal(a, b) for a, b in zip(first, path.cmd))  # the forward is deterministic (no atomics)
    for k in ("query_encoder.biLSTM.weight_hh_l0", "query_encoder.qInput.weight"):
        assert _rel_l2(g2[k], grads[k]) <= 1e-5, k


def test_query_encoder_rejects_bad_arguments(sd):
    dev = torch.device("cuda")
    path = DensePath(S.default_config(stage=1), 2, 32, dev, L=4)
    p = {k: v.to(dev) for k, v in sd.items()}
    q = path._qe_desc(p)
    q.workspace_bytes = 16
    assert L.load().drn_qe_forward(C.byref(q), L.stream_ptr()) == -1
    assert b"workspace" in L.load().drn_last_error()
    q = path._qe_desc(p)
    q.L = 65
    assert L.load().drn_qe_forward(C.byref(q), L.stream_ptr()) == -1


@pytest.mark.parametrize("M,N,K", [(32, 4096, 1024), (320, 4096, 300), (5, 70, 33), (2048, 512, 320), (32, 512, 2048)])
def test_sgemm_forms(M, N, K):
    """x W^T (+bias, relu), dy W, dy^T x with the stride conventions the path uses; accumulate and split-K paths."""
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(M + N + K)
    x, w, b = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g), torch.randn(N, generator=g)
    xd, wd, bd = x.to(dev), w.to(dev), b.to(dev)
    lib = L.load()

    def run(A, sam, sak, Bm, sbk, sbn, out, m, n, k, bias=None, relu=0, acc=0):
        L.check(lib.drn_sgemm(L.ptr(A), C.c_int64(sam), C.c_int64(sak), L.ptr(Bm), C.c_int64(sbk), C.c_int64(sbn), L.ptr(out),
                              C.c_int64(out.shape[1]), m, n, k, L.ptr(bias), relu, acc, L.stream_ptr()), "sgemm")
    ref = (x.double() @ w.double().t() + b.double())
    y = torch.full((M, N), float("nan"), device=dev)
    run(xd, K, 1, wd, 1, K, y, M, N, K, bias=bd)
    assert _rel(y, ref) <= 1e-5
    run(xd, K, 1, wd, 1, K, y, M, N, K, bias=bd, relu=1)
    assert _rel(y, ref.clamp(min=0)) <= 1e-5
    dy = torch.randn(M, N, generator=g)
    dyd = dy.to(dev)
    dx = torch.full((M, K), float("nan"), device=dev)
    run(dyd, N, 1, wd, K, 1, dx, M, K, N)
    assert _rel(dx, dy.double() @ w.double()) <= 1e-5
    dw = torch.ones(N, K, device=dev)
    run(dyd, 1, N, xd, K, 1, dw, N, K, M, acc=1)
    assert _rel(dw, 1.0 + dy.double().t() @ x.double()) <= 1e-5
    # deterministic small-batch Linear forward (fixed reduction order): exact run-to-run
    outs = []
    for _ in range(2):
        y2 = torch.full((M, N), float("nan"), device=dev)
        L.check(lib.drn_linear_fwd(L.ptr(xd), C.c_int64(K), L.ptr(wd), C.c_int64(K), L.ptr(bd), L.ptr(y2), C.c_int64(N), M, N, K, 1,
                                   L.stream_ptr()), "linear_fwd")
        outs.append(y2)
    assert _rel(outs[0], ref.clamp(min=0)) <= 1e-5 and torch.equal(outs[0], outs[1])
