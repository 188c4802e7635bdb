#!/bin/bash
# 8-GPU pass 4: the peer-memory all-reduce kernel against NCCL -- kernel check + CTA sweep, DP gradient check, phase timeline, bench A/B
set -u
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 180 $TR --master-port 29621 scripts/p2p_check.py 2>gpurun_out/r02_p2p${N}.err | grep p2p_check | tee gpurun_out/r02_p2p_check${N}.txt
DRN_EXPECT_TRANSPORT=p2p timeout 240 $TR --master-port 29622 scripts/dp_check.py 2>>gpurun_out/r02_p2p${N}.err | tail -2
for cfg in "DRN_DP_P2P=1" "DRN_DP_P2P=0"; do
  echo "---- [$cfg] bench N=$N"
  env $cfg timeout 300 $TR --master-port 29624 bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline --no-extra --sustain-seconds 0 2>>gpurun_out/r02_p2p${N}.err | tail -1 > gpurun_out/r02_bench${N}d_${cfg}.json
  python - <<PY
import json
d=json.load(open('gpurun_out/r02_bench${N}d_${cfg}.json'))
print('  value %.0f pairs/s  step %.3f ms  fwd %.3f bwd %.3f  e2e %.0f (%.3f ms)' % (d['value'], d['ms_per_step'], d['diag']['fwd_ms'], d['diag']['bwd_ms'], d['e2e']['value'], d['e2e']['ms_per_step']), d.get('gradient_exchange'))
PY
done
echo "---- timeline N=$N (peer-memory all-reduce)"
timeout 300 $TR --master-port 29623 scripts/dp_timeline.py --steps 20 2>>gpurun_out/r02_p2p${N}.err | tail -1 > gpurun_out/r02_dp_timeline${N}_p2p.json
python - gpurun_out/r02_dp_timeline${N}_p2p.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
for n,v in zip(d['phases'], d['max_over_ranks']): print('   %-42s %.3f'%(n,v))
PY
grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/r02_p2p${N}.err | tail -5
