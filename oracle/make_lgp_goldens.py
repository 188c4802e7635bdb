"""TEST INFRASTRUCTURE ONLY.  Writes tests/golden/lgp.npz by running the UNMODIFIED reference module
(/root/reference/model/LGP.py, which imports only torch) on seeded inputs: forward output, updated BatchNorm buffers, and all
gradients for an upstream gradient, in train and eval mode.

    python /root/repo/oracle/make_lgp_goldens.py        # needs /root/reference; CPU only
"""
import importlib.util
import os

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("ref_lgp", "/root/reference/model/LGP.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

out = {}
for name, (B, Cn, Q, t, training) in {"train": (5, 96, 64, 12, True), "eval": (3, 96, 64, 8, False)}.items():
    g = torch.Generator().manual_seed(222)
    m = ref.LGP(input_dim=Cn, query_dim=Q)
    with torch.no_grad():
        m.query_fc[0].weight.copy_(torch.randn(Cn, Q, 1, generator=g) * 0.2)
        m.query_fc[1].weight.copy_(torch.rand(Cn, generator=g) + 0.5)
        m.query_fc[1].bias.copy_(torch.randn(Cn, generator=g) * 0.1)
        m.query_fc[1].running_mean.copy_(torch.randn(Cn, generator=g) * 0.1)
        m.query_fc[1].running_var.copy_(torch.rand(Cn, generator=g) + 0.5)
    m.train(training)
    x = torch.randn(B, Cn, t, generator=g, requires_grad=True)
    q = torch.randn(B, Q, generator=g, requires_grad=True)
    dout = torch.randn(B, Cn, t // 2, generator=g)
    for k, v in m.state_dict().items():
        out["%s/init/%s" % (name, k)] = v.numpy().copy()
    y = m(x, q)
    y.backward(dout)
    out.update({name + "/x": x.detach().numpy(), name + "/q": q.detach().numpy(), name + "/dout": dout.numpy(), name + "/y": y.detach().numpy(),
                name + "/dx": x.grad.numpy(), name + "/dq": q.grad.numpy()})
    for k, v in m.named_parameters():
        out["%s/grad/%s" % (name, k)] = v.grad.numpy()
    for k, v in m.state_dict().items():
        out["%s/after/%s" % (name, k)] = v.numpy().copy()
np.savez_compressed(os.path.join(REPO, "tests", "golden", "lgp.npz"), **out)
print("wrote", len(out), "arrays")
