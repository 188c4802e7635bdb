#!/bin/bash
# N = 2 and N = 4 with the peer-memory all-reduce (compare with the NCCL runs of the same round)
set -u
mkdir -p gpurun_out
for N in 2 4; do
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for cfg in "DRN_DP_P2P=1" "DRN_DP_P2P=0"; do
  echo "---- [$cfg] bench N=$N"
  env $cfg timeout 300 $TR --master-port 2963$N bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline --no-extra --sustain-seconds 0 2>>gpurun_out/r02_p2p24.err | tail -1 > gpurun_out/r02_bench${N}e_${cfg}.json
  python - <<PY
import json
d=json.load(open('gpurun_out/r02_bench${N}e_${cfg}.json'))
print('  value %.0f pairs/s  step %.3f ms  fwd %.3f bwd %.3f  e2e %.0f (%.3f ms)' % (d['value'], d['ms_per_step'], d['diag']['fwd_ms'], d['diag']['bwd_ms'], d['e2e']['value'], d['e2e']['ms_per_step']), d.get('gradient_exchange'))
PY
done
done
grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/r02_p2p24.err | tail -5
