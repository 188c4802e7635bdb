#!/bin/bash
# peer-memory all-reduce bring-up on N GPUs (default 2): kernel check, DP gradient check on both transports, bench A/B
set -u
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 180 $TR --master-port 29611 scripts/p2p_check.py 2>&1 | grep -v "^\*\*\*\|OMP_NUM" | tail -15
echo "p2p_check rc=$?"
DRN_EXPECT_TRANSPORT=p2p timeout 240 $TR --master-port 29612 scripts/dp_check.py 2>&1 | grep -v "^\*\*\*\|OMP_NUM" | tail -6
DRN_DP_P2P=0 DRN_EXPECT_TRANSPORT=nccl timeout 240 $TR --master-port 29613 scripts/dp_check.py 2>&1 | grep -v "^\*\*\*\|OMP_NUM" | tail -3
for cfg in "DRN_DP_P2P=1" "DRN_DP_P2P=0"; do
  echo "---- [$cfg] bench N=$N"
  env $cfg timeout 300 $TR --master-port 29614 bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline --no-extra --sustain-seconds 0 2> gpurun_out/r02_bench${N}_p2p.err | tail -1 > gpurun_out/r02_bench${N}_${cfg}.json
  python - <<PY
import json
d=json.load(open('gpurun_out/r02_bench${N}_${cfg}.json'))
print('  value %.0f pairs/s  step %.3f ms  fwd %.3f bwd %.3f  e2e %.0f (%.3f ms)' % (d['value'], d['ms_per_step'], d['diag']['fwd_ms'], d['diag']['bwd_ms'], d['e2e']['value'], d['e2e']['ms_per_step']), d.get('gradient_exchange'))
PY
done
tail -5 gpurun_out/r02_bench${N}_p2p.err
