"""Probe: do the query-encoder kernels actually run beside the persistent CTA-pair GEMM (SM co-residency)?  Times the prop_fc
weight gradient alone, the query-encoder backward alone, and both on two streams."""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from drn_b200 import lib as L, ops, spec as spec_mod, synthetic as S
from model.main_model import mainModel

cfg = S.default_config(stage=1)
sd = S.synth_state_dict(spec_mod.state_dict_spec(cfg))
batch = S.synth_batch(32, 256, max_len=10, embedding=sd["query_encoder.embedding.weight"])
model = mainModel(1301, S.config_namespace(stage=1)); model.load_state_dict(sd)
for k, p in model.named_parameters():
    if "iou_scores" in k or "mix_fc" in k: p.requires_grad = False
model = model.cuda().train()
for _ in range(2):
    for p in model.parameters(): p.grad = None
    _, ld = model(batch["query_tokens"], batch["query_length"], batch["props_features"], batch["props_start_end"], batch["gt_start_end"], None, None)
    sum(ld.values()).backward()
torch.cuda.synchronize()
path = list(model._paths.values())[0]
p = model._tensor_dict()
names = model._trainable_names
grads = {n: torch.zeros_like(p[n]) for n in names}
side = torch.cuda.Stream()
lib = L.load()

def gemm():
    ops.gemm(L.GEMM_WGRAD, path.dP_pl.desc(), path.f_pl.desc(), 32, 256, 4096, M=4096, out=grads["prop_fc.weight"], out_ld=4096, out_tap_stride=0)
def qe_b():
    L.check(lib.drn_qe_backward(C.byref(path._qe_desc(p, grads)), L.stream_ptr()), "qe_b")
def qe_f():
    L.check(lib.drn_qe_forward(C.byref(path._qe_desc(p)), L.stream_ptr()), "qe_f")

def timeit(fn, n=10):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

def both(q):
    def f():
        ev = torch.cuda.Event(); ev.record(); side.wait_event(ev)
        with torch.cuda.stream(side): gemm()
        q()
        ev2 = torch.cuda.Event(); ev2.record(side); torch.cuda.current_stream().wait_event(ev2)
    return f

print("gemm alone      %.3f ms" % timeit(gemm))
print("qe_bwd alone    %.3f ms" % timeit(qe_b))
print("qe_fwd alone    %.3f ms" % timeit(qe_f))
print("gemm || qe_bwd  %.3f ms" % timeit(both(qe_b)))
print("gemm || qe_fwd  %.3f ms" % timeit(both(qe_f)))

# sanity: two independent chains on two streams, and the GEMM beside a plain elementwise kernel
big = torch.zeros(64 << 20, device="cuda")
def ew():
    for _ in range(20): big.add_(1.0)
def two_qe():
    ev = torch.cuda.Event(); ev.record(); side.wait_event(ev)
    with torch.cuda.stream(side): qe_f()
    qe_f()
    ev2 = torch.cuda.Event(); ev2.record(side); torch.cuda.current_stream().wait_event(ev2)
print("ew alone        %.3f ms" % timeit(ew))
print("gemm || ew      %.3f ms" % timeit(both(ew)))
print("qe_fwd || qe_fwd(same buffers) %.3f ms" % timeit(two_qe))
