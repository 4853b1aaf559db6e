#!/bin/bash
# Launch list of one bs=1 forward (eager, no graph / side stream) -> gpurun_out/launches_bs1.csv
mkdir -p gpurun_out
export VLA_NO_GRAPH=1 VLA_NO_SIDE_STREAM=1
CMD="python bench.py --batch 1 --steps 1 --warmup 3 --no-cpu-baseline --latency-iters 0"
$CMD > gpurun_out/plain_bs1.log 2>&1 || { tail -5 gpurun_out/plain_bs1.log; exit 1; }
N=$(python -c "import json;print(json.loads(open('gpurun_out/plain_bs1.log').read().strip().splitlines()[-1])['gpu_launches'])")
echo "launches per step: $N"
KRE='gemm_bf16|flash_attn|splitkv_attn|fa_tcgen05|norm_kernel|rope_apply|im2col|prefix_tokens|assemble|skinny|policy_|head_out|broadcast_row|gather_rows|copy_view'
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$KRE" -s $((3*N)) -c $N --csv --log-file gpurun_out/launches_bs1.csv $CMD > gpurun_out/ncu_bs1.log 2>&1
echo "rc=$?"
