"""drn_clip_adam (fused clip_grad_norm_ + Adam, reference main.py:239-244) against torch.nn.utils.clip_grad_norm_ +
torch.optim.Adam on the same parameters and gradients: several steps, clipping active and inactive, parameters that only
enter the norm (stage-2 quirk, main.py:124-138), a parameter without gradient.  Floating point: parameters within 1e-6
relative of torch's after every step (same formula, different summation order of the norm)."""
import pytest
import torch

from drn_b200.optim import FusedClipAdam

pytestmark = pytest.mark.gpu


def _params(seed):
    g = torch.Generator().manual_seed(seed)
    shapes = [(4096, 300), (1, ), (256, 4352, 3), (17, ), (512, 512, 3), (3, 5, 7)]
    return [torch.randn(*s, generator=g).cuda().requires_grad_(True) for s in shapes]


@pytest.mark.parametrize("max_norm", [0.5, 1e9])
def test_matches_torch_clip_and_adam(max_norm):
    ours, ref = _params(0), _params(0)
    upd = [0, 1, 2, 4]          # parameters 3 and 5 are outside the optimizer: they only inflate the clipping norm
    opt = FusedClipAdam([ours[i] for i in upd], lr=1e-3, clip_params=ours, max_norm=max_norm)
    topt = torch.optim.Adam([ref[i] for i in upd], lr=1e-3)
    g = torch.Generator().manual_seed(1)
    for step in range(4):
        for i, (a, b) in enumerate(zip(ours, ref)):
            if i == 1 and step == 2:   # a parameter whose gradient is missing this step
                a.grad = b.grad = None
                continue
            gr = (torch.randn(a.shape, generator=g) * (10.0 if step % 2 else 0.01)).cuda()
            a.grad, b.grad = gr.clone(), gr.clone()
        tn = torch.nn.utils.clip_grad_norm_(ref, max_norm)
        topt.step()
        opt.step()
        torch.cuda.synchronize()
        assert abs(float(opt.total_norm()) - float(tn)) <= 1e-5 * float(tn)
        for a, b in zip(ours, ref):
            assert float((a - b).abs().max()) <= 1e-6 * max(float(b.abs().max()), 1.0), step
    # untouched: parameters outside the optimizer
    p0 = _params(0)
    assert torch.equal(ours[3], p0[3]) and torch.equal(ours[5], p0[5])
