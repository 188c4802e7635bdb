#!/bin/bash
# A/B on ONE box: bash scripts/ab_bench.sh "ENV_A=1" "ENV_B=2" ...  (each argument = environment for one bench run; "" = default)
for cfg in "$@"; do
  echo "---- [$cfg]"
  for rep in 1 2; do
    env $cfg timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extra --sustain-seconds 0 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('  step %.3f ms  fwd %.3f  bwd %.3f  e2e %.3f' % (d['ms_per_step'], d['diag']['fwd_ms'], d['diag']['bwd_ms'], d['e2e']['ms_per_step']), 'host enqueue', d['diag'].get('host_enqueue_ms_per_step'))"
  done
done
