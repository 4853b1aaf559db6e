#!/bin/bash
# Launch list of one bench step (own kernels only) + per-GEMM-shape timing CSV.
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --latency-iters 0"
rm -f gpurun_out/gemm_shapes.csv
VLA_GEMM_PROF_CSV=gpurun_out/gemm_shapes.csv $CMD > gpurun_out/plain.log 2>&1 || { tail -5 gpurun_out/plain.log; exit 1; }
N=$(python -c "import json;print(json.loads(open('gpurun_out/plain.log').read().strip().splitlines()[-1])['gpu_launches'])")
echo "launches per step: $N"
ncu --metrics gpu__time_duration.sum --clock-control none \
    -k regex:'gemm_bf16|flash_attn|norm_kernel|rope_apply|im2col|prefix_tokens|assemble|skinny|policy_|head_out|broadcast_row|gather_rows|copy_view|attn' \
    -s $((3*N)) -c $N --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu.log 2>&1
echo "rc=$?"
tail -3 gpurun_out/plain.log | cut -c1-400
wc -l gpurun_out/launches.csv gpurun_out/gemm_shapes.csv
