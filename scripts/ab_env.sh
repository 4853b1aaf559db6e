#!/bin/bash
# A/B of two environments on the SAME box: scripts/ab_env.sh "ENV_A=.. ENV_B=.." "ENV_C=.." [bench args]  (A B A B)
A=$1; B=$2; shift 2
for envs in "$A" "$B" "$A" "$B"; do
  env $envs python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline --latency-iters 100 "$@" 2>/dev/null | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); s=d['segments']; print('[$envs]', 'value %.1f ms %.2f gemm %.2f frac %.3f towers %.2f llm %.2f policy %.2f bs1 %.3f launches/step %d' % (d['value'], d['ms_per_step'], d['roofline']['gemm_ms_per_step'], d['roofline']['frac'], s['towers_projector_ms'], s['llm_prefill_ms'], s['policy_ms'], d['latency_bs1']['p50_ms'], d['gpu_launches']/10))"
done
