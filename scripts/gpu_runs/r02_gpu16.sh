#!/bin/bash
# 8-GPU pass 3: 'split' order (head+FPN gradients all-reduced during the backbone backward) against tail_first
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
DRN_DP_ORDER=split timeout 300 $TR --master-port 29560 scripts/dp_check.py 2>gpurun_out/r02_dpcheck8.err | tail -2
for cfg in "DRN_DP_ORDER=split" "DRN_DP_ORDER=tail_first"; do
  echo "---- [$cfg] timeline N=8"
  env $cfg timeout 300 $TR --master-port 29561 scripts/dp_timeline.py --steps 20 2>>gpurun_out/r02_tl8c.err | tail -1 > gpurun_out/r02_dp_timeline8_${cfg#*=}.json
  python - gpurun_out/r02_dp_timeline8_${cfg#*=}.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
for n,v in zip(d['phases'], d['max_over_ranks']): print('   %-42s %.3f'%(n,v))
PY
  echo "---- [$cfg] bench N=8"
  env $cfg timeout 600 $TR --master-port 29562 bench.py --gpus 8 --steps 30 --warmup 5 --no-extra --sustain-seconds 0 2>gpurun_out/r02_bench8c.err | tail -1 > gpurun_out/r02_bench8c_${cfg#*=}.json
  python -c "import json,sys; d=json.load(open(sys.argv[1])); print('  value %.0f pairs/s  step %.3f ms  fwd %.3f bwd %.3f  e2e %.0f (%.3f ms)'%(d['value'],d['ms_per_step'],d['diag']['fwd_ms'],d['diag']['bwd_ms'],d['e2e']['value'],d['e2e']['ms_per_step']))" gpurun_out/r02_bench8c_${cfg#*=}.json
done
tail -3 gpurun_out/r02_tl8c.err
