#!/bin/bash
# Profiles of one bench step (own kernels only), each taken after the plain command exited 0:
#   1. launch list: per-kernel gpu__time_duration (+ DRAM bytes) for every launch of one step -> launches.csv
#   2. per-GEMM-shape CUDA-event timings -> gemm_shapes.csv
#   3. one `ncu --set full` capture of the dominant GEMM launch and of each tcgen05 attention shape -> *.ncu-rep
# CUDA-graph replay is switched off (VLA_NO_GRAPH=1) so that ncu's -s/-c launch counting sees plain launches.
mkdir -p gpurun_out
export VLA_NO_GRAPH=1
export VLA_BENCH_STAGE_LIMIT_S=1500   # ncu replays every kernel: a stage takes minutes
CMD="timeout 900 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline --latency-iters 0"
rm -f gpurun_out/gemm_shapes.csv
VLA_GEMM_PROF_CSV=gpurun_out/gemm_shapes.csv $CMD > gpurun_out/plain.log 2>&1 || { tail -5 gpurun_out/plain.log; exit 1; }
N=$(python -c "import json;print(json.loads([l for l in open('gpurun_out/plain.log') if l.startswith('{')][-1])['gpu_launches'])")
echo "launches per step: $N"
KRE='gemm_bf16|flash_attn|splitkv_attn|fa_tcgen05|norm_kernel|row_stats|rope_apply|im2col|prefix_tokens|assemble|skinny|policy_|head_out|broadcast_row|gather_rows|copy_view'
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -k regex:"$KRE" -s $((3*N)) -c $N --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu.log 2>&1
echo "launch list rc=$?"
tail -2 gpurun_out/plain.log | cut -c1-300
wc -l gpurun_out/launches.csv gpurun_out/gemm_shapes.csv
[ -n "$SKIP_FULL" ] && exit 0   # launch list only
NG=$(wc -l < gpurun_out/gemm_shapes.csv)   # GEMM launches per step
# full captures: the first LLM gate/up GEMM of the 4th step (148 CTAs = 74 pairs, SwiGLU epilogue) and one attention launch per shape
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s $((3*NG+203)) -c 1 -f -o gpurun_out/gemm_full $CMD > gpurun_out/ncu_gemm_full.log 2>&1
echo "gemm full rc=$?"
ncu --set full --clock-control none --import-source on -k regex:fa_tcgen05 -s $((3*73+2)) -c 1 -f -o gpurun_out/attn_full_dino $CMD > gpurun_out/ncu_attn_full.log 2>&1
echo "attn full rc=$?"
ls -la gpurun_out/*.ncu-rep
