#!/bin/bash
# 8-GPU pass 2: overlap order (B+C reduced beside the tail, NCCL capped at 16 CTAs) against tail_first; bench of the best
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
i=10
for cfg in "DRN_DP_ORDER=overlap" "DRN_DP_ORDER=overlap DRN_NCCL_MAX_CTAS=default" "DRN_DP_ORDER=tail_first" "DRN_DP_ORDER=overlap DRN_NCCL_MAX_CTAS=12"; do
  i=$((i+1))
  echo "---- [$cfg] timeline N=8"
  env $cfg timeout 300 $TR --master-port $((29520+i)) scripts/dp_timeline.py --steps 20 2>>gpurun_out/r02_tl8b.err | tail -1 > gpurun_out/r02_dp_timeline8_$i.json
  python - gpurun_out/r02_dp_timeline8_$i.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print('  nccl_max_ctas', d['nccl_max_ctas'], 'chunks', d['chunks'])
for n,v in zip(d['phases'], d['max_over_ranks']): print('   %-42s %.3f'%(n,v))
PY
done
for cfg in "DRN_DP_ORDER=overlap" "DRN_DP_ORDER=tail_first"; do
  echo "---- [$cfg] bench N=8"
  env $cfg timeout 600 $TR --master-port 29541 bench.py --gpus 8 --steps 30 --warmup 5 --no-extra --sustain-seconds 0 2>gpurun_out/r02_bench8b.err | tail -1 > gpurun_out/r02_bench8b_${cfg#*=}.json
  python -c "import json,sys; d=json.load(open(sys.argv[1])); print('  value %.0f pairs/s  step %.3f ms  fwd %.3f bwd %.3f  e2e %.0f (%.3f ms)'%(d['value'],d['ms_per_step'],d['diag']['fwd_ms'],d['diag']['bwd_ms'],d['e2e']['value'],d['e2e']['ms_per_step']))" gpurun_out/r02_bench8b_${cfg#*=}.json
done
tail -3 gpurun_out/r02_tl8b.err
