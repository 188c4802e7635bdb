#!/bin/bash
# 8-GPU confirmation of the final code: DP gradient check + bench on the peer-memory all-reduce
set -u
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
DRN_EXPECT_TRANSPORT=p2p timeout 240 $TR --master-port 29642 scripts/dp_check.py 2>gpurun_out/r02_final${N}.err | tail -2
timeout 300 $TR --master-port 29644 bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline --no-extra --sustain-seconds 0 2>>gpurun_out/r02_final${N}.err | tail -1 > gpurun_out/r02_bench${N}_final.json
python - <<PY
import json
d=json.load(open('gpurun_out/r02_bench${N}_final.json'))
print('  value %.0f pairs/s  step %.3f ms  fwd %.3f bwd %.3f  e2e %.0f (%.3f ms)' % (d['value'], d['ms_per_step'], d['diag']['fwd_ms'], d['diag']['bwd_ms'], d['e2e']['value'], d['e2e']['ms_per_step']), d.get('gradient_exchange'))
PY
grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/r02_final${N}.err | tail -3
