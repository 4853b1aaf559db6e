#!/bin/bash
# A/B of one environment switch on the same box: scripts/ab_env.sh VAR [rounds]  (bench value with VAR unset / VAR=1)
VAR=$1; R=${2:-2}
for i in $(seq $R); do
  for v in "" 1; do
    if [ -z "$v" ]; then unset $VAR; else export $VAR=$v; fi
    timeout 240 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --latency-iters 0 2>/dev/null | python -c "
import json,sys,os
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
s=d['segments']
print('$VAR=%s' % os.environ.get('$VAR',''), round(d['value'],1), round(d['ms_per_step'],2), 'towers', round(s['towers_projector_ms'],2), 'prefill', round(s['llm_prefill_ms'],2), 'policy', round(s['policy_ms'],2), 'gemm_ms', round(d['roofline']['gemm_ms_per_step'],2), d['clocks']['sm_mhz'])
"
  done
done
