#!/bin/bash
# Operator-level GPU check: smallest GEMM first (a deadlock there must not eat the budget), then all ops, then GEMM TF/s.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 180 python -m pytest tests/test_ops_gpu.py -q -x -k "gemm_plain and 128-256-64" > gpurun_out/ops_first.log 2>&1
rc=$?
tail -15 gpurun_out/ops_first.log
if [ $rc -ne 0 ]; then echo "FIRST GEMM FAILED rc=$rc"; exit $rc; fi
timeout 900 python -m pytest tests/test_ops_gpu.py -q > gpurun_out/ops_all.log 2>&1
echo "ops_all rc=$?"
tail -60 gpurun_out/ops_all.log
timeout 300 python tests/gemm_bench_gpu.py > gpurun_out/gemm_bench.log 2>&1
echo "bench rc=$?"
cat gpurun_out/gemm_bench.log
