#!/bin/bash
# 2-GPU pass: DP parity test (3 modes), whole GPU suite, N=2 bench for both backward orders, phase timelines
set -u
mkdir -p gpurun_out
nvidia-smi -L
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r02_pytest_gpu_e.log
grep -E "passed|failed|^E  |FAILED" gpurun_out/r02_pytest_gpu_e.log | head
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for cfg in "DRN_DP_ORDER=r01" "DRN_DP_ORDER=tail_first"; do
  echo "---- [$cfg] bench N=2"
  env $cfg timeout 600 $TR --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 --no-extra --sustain-seconds 0 2>gpurun_out/r02_bench2.err | tail -1 > gpurun_out/r02_bench2_${cfg#*=}.json
  python -c "import json,sys; d=json.load(open(sys.argv[1])); print('  value %.0f pairs/s  step %.3f ms  fwd %.3f bwd %.3f  e2e %.0f (%.3f ms) h2d %s'%(d['value'],d['ms_per_step'],d['diag']['fwd_ms'],d['diag']['bwd_ms'],d['e2e']['value'],d['e2e']['ms_per_step'],d['e2e']['h2d_gbs_per_rank_all_ranks_uploading']))" gpurun_out/r02_bench2_${cfg#*=}.json
  env $cfg timeout 300 $TR --master-port 29512 scripts/dp_timeline.py --steps 20 2>>gpurun_out/r02_bench2.err | tail -1 > gpurun_out/r02_dp_timeline2_${cfg#*=}.json
  cat gpurun_out/r02_dp_timeline2_${cfg#*=}.json | cut -c1-1500
done
timeout 300 python scripts/dp_timeline.py --steps 20 2>>gpurun_out/r02_bench2.err | tail -1 > gpurun_out/r02_dp_timeline1.json
cat gpurun_out/r02_dp_timeline1.json | cut -c1-800
tail -5 gpurun_out/r02_bench2.err
