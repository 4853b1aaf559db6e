#!/bin/bash
# A/B of two builds of the library on the SAME box: scripts/ab_bench.sh libA.so libB.so [extra bench args]
# prints value / ms_per_step / gemm ms / segments / bs=1 p50 per run, A B A B
A=$1; B=$2; shift 2
for lib in $A $B $A $B; do
  VLA_B200_LIB=$PWD/vla_adapter_b200/lib/$lib python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline --latency-iters 100 "$@" 2>/dev/null | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); s=d['segments']; print('$lib', 'value %.1f ms %.2f gemm %.2f towers %.2f llm %.2f policy %.2f bs1 %.3f clk %s' % (d['value'], d['ms_per_step'], d['roofline']['gemm_ms_per_step'], s['towers_projector_ms'], s['llm_prefill_ms'], s['policy_ms'], d['latency_bs1']['p50_ms'], d['clocks']['sm_mhz']))"
done
