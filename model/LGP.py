"""`LGP`: language-guided pooling, same interface as the reference model/LGP.py:4-51 (constructor, `query_fc` parameter names,
`forward(inputs [B, C, t], query [B, Q]) -> [B, C, t/2]`).  Dead code in the reference (imported by nothing) but named in
the task's north star as part of the query-video fusion; provided as a standalone fused op.  The math
runs in libdrn_sm100 (drn_b200/csrc/lgp.cu + the small-Linear kernels); there is no torch fallback."""
import ctypes as C

import torch
import torch.nn as nn

from drn_b200 import lib as L


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class _LGPFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, query, weight, gamma, beta, running_mean, running_var, nbt, training, momentum, eps):
        lib, st = L.load(), L.stream_ptr()
        x, query = x.contiguous().float(), query.contiguous().float()
        B, Cn, t = x.shape
        Q = query.shape[1]
        w2 = weight.view(Cn, Q)
        z = torch.empty(B, Cn, device=x.device)
        # z = query W^T in exact fp32 (store mode: no split, no atomics); the batch statistics of a handful of query vectors
        # that follow amplify rounding noise, so this one does not go through the split-BF16 drn_linear_fwd
        fj = (L.SgemmJob * 1)()
        f = fj[0]
        f.A, f.sam, f.sak, f.B, f.sbk, f.sbn, f.C, f.ldc, f.M, f.N, f.K = query.data_ptr(), Q, 1, w2.data_ptr(), 1, Q, z.data_ptr(), Cn, B, Cn, Q
        f.store = 1
        L.check(lib.drn_sgemm_batch(1, fj, st), "lgp linear")
        qn, xhat, invstd = torch.empty_like(z), torch.empty_like(z), torch.empty(Cn, device=x.device)
        L.check(lib.drn_lgp_bn(_p(z), B, Cn, t, _p(gamma), _p(beta), _p(running_mean), _p(running_var), _p(nbt), C.c_float(momentum),
                               C.c_float(eps), 1 if training else 0, _p(qn), _p(xhat), _p(invstd), st), "lgp_bn")
        att = torch.empty(B, t // 2, 2, device=x.device)
        out = torch.empty(B, Cn, t // 2, device=x.device)
        L.check(lib.drn_lgp_pool_fwd(_p(x), _p(qn), B, Cn, t, _p(att), _p(out), st), "lgp_pool_fwd")
        ctx.save_for_backward(x, query, w2, gamma, qn, xhat, invstd, att)
        ctx.training = training
        return out

    @staticmethod
    def backward(ctx, dout):
        lib, st = L.load(), L.stream_ptr()
        x, query, w2, gamma, qn, xhat, invstd, att = ctx.saved_tensors
        B, Cn, t = x.shape
        Q = query.shape[1]
        dout = dout.contiguous().float()
        dx, dqn = torch.empty_like(x), torch.empty(B, Cn, device=x.device)
        L.check(lib.drn_lgp_pool_bwd(_p(x), _p(qn), _p(att), _p(dout), B, Cn, t, _p(dx), _p(dqn), st), "lgp_pool_bwd")
        dz = torch.empty(B, Cn, device=x.device)
        dgamma, dbeta = torch.zeros_like(gamma), torch.zeros_like(gamma)
        L.check(lib.drn_lgp_bn_bwd(_p(dqn), _p(xhat), _p(invstd), _p(gamma), B, Cn, 1 if ctx.training else 0, _p(dz), _p(dgamma),
                                   _p(dbeta), st), "lgp_bn_bwd")
        dquery, dw = torch.zeros(B, Q, device=x.device), torch.zeros(Cn, Q, device=x.device)
        jobs = (L.SgemmJob * 2)()
        a, b = jobs[0], jobs[1]  # d query = dz W ; dW = dz^T query
        a.A, a.sam, a.sak, a.B, a.sbk, a.sbn, a.C, a.ldc, a.M, a.N, a.K = dz.data_ptr(), Cn, 1, w2.data_ptr(), Q, 1, dquery.data_ptr(), Q, B, Q, Cn
        b.A, b.sam, b.sak, b.B, b.sbk, b.sbn, b.C, b.ldc, b.M, b.N, b.K = dz.data_ptr(), 1, Cn, query.data_ptr(), Q, 1, dw.data_ptr(), Q, Cn, Q, B
        L.check(lib.drn_sgemm_batch(2, jobs, st), "lgp sgemm")
        return dx, dquery, dw.view(Cn, Q, 1), dgamma, dbeta, None, None, None, None, None, None


class LGP(nn.Module):
    def __init__(self, input_dim=1024, query_dim=1024, use_bn=True):
        super().__init__()
        if not use_bn:
            raise NotImplementedError("the B200 LGP op implements the reference's default (use_bn=True)")
        conv = nn.Conv1d(query_dim, input_dim, kernel_size=1, stride=1, padding=0, dilation=1, bias=False)
        nn.init.kaiming_uniform_(conv.weight, a=1)
        self.query_fc = nn.Sequential(conv, nn.BatchNorm1d(input_dim))

    def forward(self, inputs, query):
        if inputs.device.type != "cuda":
            raise RuntimeError("LGP runs on a B200 through libdrn_sm100.so only (no CPU / torch fallback)")
        if inputs.size(-1) % 2:
            raise ValueError("LGP pools pairs of time steps: t must be even (reference LGP.py:45 views [.., t/2, 2])")
        conv, bn = self.query_fc[0], self.query_fc[1]
        return _LGPFn.apply(inputs, query, conv.weight, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.num_batches_tracked,
                            self.training, bn.momentum, bn.eps)
