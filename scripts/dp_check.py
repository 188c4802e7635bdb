"""Data-parallel correctness on real GPUs (run under torchrun, >= 2 ranks): the gradients DataParallelDRN leaves in param.grad
(two-part backward, all-reduce overlapped with the tail) must equal the mean over ranks of the single-rank gradients of each
rank's own shard (SURVEY.md section 8e parity oracle for DP).  Prints max relative deviation per rank-0."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from drn_b200 import spec as spec_mod  # noqa: E402
from drn_b200 import synthetic as S  # noqa: E402
from drn_b200.parallel import DataParallelDRN  # noqa: E402
from model.main_model import mainModel  # noqa: E402

AUTO = "--auto" in sys.argv  # no wrapper, no init_process_group: mainModel's WORLD_SIZE hook must do both (torchrun main.py)
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if not AUTO:
    from drn_b200.parallel import nccl_env_defaults
    nccl_env_defaults()
    dist.init_process_group("nccl", device_id=dev)
cfg = S.default_config(stage=1)
sd = S.synth_state_dict(spec_mod.state_dict_spec(cfg))
B, T = 8, 64
batch = S.synth_batch(B, T, max_len=10, embedding=sd["query_encoder.embedding.weight"], seed=S.SEED + 17 * rank)


def make():
    m = mainModel(1301, S.config_namespace(stage=1))
    m.load_state_dict(sd)
    for k, p in m.named_parameters():
        if "iou_scores" in k or "mix_fc" in k:
            p.requires_grad = False
    return m.to(dev).train()


def step(m):
    for p in m.parameters():
        p.grad = None
    _, ld = m(batch["query_tokens"], batch["query_length"], batch["props_features"], batch["props_start_end"], batch["gt_start_end"], None, None)
    sum(ld.values()).backward()
    torch.cuda.synchronize()
    return {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}


os.environ["DRN_AUTO_DP"] = "0"  # the single-rank reference gradients come from an unsynchronised replica
local_model = make()
g_local = step(local_model)
os.environ["DRN_AUTO_DP"] = "1"
if AUTO:
    dp_model = make()
    with torch.no_grad():  # ranks start from different weights: the hook must broadcast rank 0's
        dp_model.prop_fc.bias.add_(float(rank))
    step(dp_model)  # first forward: creates the process group, broadcasts, installs the reducer
    assert dist.is_initialized() and dp_model._dp is not None and dp_model._dp.world == world
    t = dp_model.prop_fc.bias.detach().clone()
    dist.broadcast(t, 0)
    assert torch.equal(t, dp_model.prop_fc.bias.detach()), "rank-0 broadcast missing"
    with torch.no_grad():
        dp_model.prop_fc.bias.copy_(sd["prop_fc.bias"])
expected = {}
for k, g in g_local.items():
    t = g.clone()
    dist.all_reduce(t)
    expected[k] = t / world
dp_module = dp_model if AUTO else DataParallelDRN(make()).module
worst = 0.0
for it in range(3):  # eager first call, then the graph replay path twice
    g_dp = step(dp_module)
    for k, e in expected.items():
        n = float(e.norm())
        if n < 1e-6:
            continue
        worst = max(worst, float((g_dp[k] - e).norm()) / n)
t = torch.tensor([worst], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
red = dp_module._dp
want = os.environ.get("DRN_EXPECT_TRANSPORT")
if want:
    assert red.transport_used[want] > 0 and red.transport_used["p2p" if want == "nccl" else "nccl"] == 0, red.transport_used
if rank == 0:
    print("transport", red.transport_used)
    print("dp_check%s order=%s world=%d max rel-L2 deviation of DP gradients from the mean of single-rank gradients: %.3e" % (" (auto hook)" if AUTO else "", os.environ.get("DRN_DP_ORDER", "tail_first"), world, float(t)))
    assert float(t) < 1e-3
dist.destroy_process_group()
