"""`model.LGP.LGP` (language-guided pooling op, reference model/LGP.py:29-51) against fixtures produced by the UNMODIFIED
reference module (oracle/make_lgp_goldens.py): forward, BatchNorm buffer updates, and every gradient, train and eval mode.
Floating point: 1e-5 relative (max-norm) forward, 1e-4 gradients (exact fp32 arithmetic on both sides, different summation
order)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double()
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-30))


@pytest.mark.parametrize("name", ["train", "eval"])
def test_lgp_matches_reference_golden(name, golden_dir):
    from model.LGP import LGP
    g = np.load(os.path.join(golden_dir, "lgp.npz"))
    x, q, dout = (torch.from_numpy(g["%s/%s" % (name, k)]).cuda() for k in ("x", "q", "dout"))
    Cn, Q = x.shape[1], q.shape[1]
    m = LGP(input_dim=Cn, query_dim=Q)
    m.load_state_dict({k[len(name) + 6:]: torch.from_numpy(g[k]) for k in g.files if k.startswith(name + "/init/")})
    m = m.cuda().train(name == "train")
    x.requires_grad_(True)
    q.requires_grad_(True)
    y = m(x, q)
    y.backward(dout)
    torch.cuda.synchronize()
    assert y.shape == g[name + "/y"].shape and _rel(y.detach(), g[name + "/y"]) <= 1e-5
    assert _rel(x.grad, g[name + "/dx"]) <= 1e-4 and _rel(q.grad, g[name + "/dq"]) <= 1e-4
    for k, v in m.named_parameters():
        assert _rel(v.grad, g["%s/grad/%s" % (name, k)]) <= 1e-4, k
    for k, v in m.state_dict().items():
        ref = g["%s/after/%s" % (name, k)]
        if v.is_floating_point():
            assert _rel(v, ref) <= 1e-5, k
        else:
            assert int(v) == int(ref), k


def test_lgp_rejects_odd_t_and_cpu():
    from model.LGP import LGP
    m = LGP(16, 8)
    with pytest.raises(RuntimeError, match="no CPU"):
        m(torch.zeros(1, 16, 4), torch.zeros(1, 8))
    with pytest.raises(ValueError):
        m.cuda()(torch.zeros(1, 16, 5, device="cuda"), torch.zeros(1, 8, device="cuda"))
