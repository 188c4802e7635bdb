#!/bin/bash
# 8-GPU pass: phase timelines of the DP backward for both orders and NCCL CTA caps, bench for both orders
set -u
mkdir -p gpurun_out
nvidia-smi -L | wc -l
nvidia-smi topo -m > gpurun_out/r02_topo8.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
i=0
for cfg in "DRN_DP_ORDER=r01" "DRN_DP_ORDER=tail_first" "DRN_DP_ORDER=r01 DRN_NCCL_MAX_CTAS=8" "DRN_DP_ORDER=tail_first DRN_NCCL_MAX_CTAS=8" "DRN_DP_ORDER=tail_first DRN_NCCL_MAX_CTAS=16 DRN_DP_CHUNKS=2" "DRN_DP_ORDER=r01 DRN_DP_PAIR_CLUSTERS=64"; do
  i=$((i+1))
  echo "---- [$cfg] timeline N=8"
  env $cfg NCCL_DEBUG=WARN timeout 300 $TR --master-port $((29520+i)) scripts/dp_timeline.py --steps 20 2>>gpurun_out/r02_tl8.err | tail -1 > gpurun_out/r02_dp_timeline8_$i.json
  python - gpurun_out/r02_dp_timeline8_$i.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print('  nccl_max_ctas', d['nccl_max_ctas'], 'chunks', d['chunks'])
for n,v in zip(d['phases'], d['max_over_ranks']): print('   %-42s %.3f'%(n,v))
print('   fwd+bwd max-over-ranks sum: %.3f ms'%(d['max_over_ranks'][0]+d['max_over_ranks'][-1]))
PY
done
for cfg in "DRN_DP_ORDER=r01" "DRN_DP_ORDER=tail_first"; do
  echo "---- [$cfg] bench N=8"
  env $cfg timeout 600 $TR --master-port 29540 bench.py --gpus 8 --steps 30 --warmup 5 --no-extra --sustain-seconds 0 2>gpurun_out/r02_bench8.err | tail -1 > gpurun_out/r02_bench8_${cfg#*=}.json
  python -c "import json,sys; d=json.load(open(sys.argv[1])); print('  value %.0f pairs/s  step %.3f ms  fwd %.3f bwd %.3f  e2e %.0f (%.3f ms) h2d %s aff %s'%(d['value'],d['ms_per_step'],d['diag']['fwd_ms'],d['diag']['bwd_ms'],d['e2e']['value'],d['e2e']['ms_per_step'],d['e2e']['h2d_gbs_per_rank_all_ranks_uploading'],d['e2e']['host_affinity_rank0']))" gpurun_out/r02_bench8_${cfg#*=}.json
done
tail -3 gpurun_out/r02_tl8.err
