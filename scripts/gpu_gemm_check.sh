#!/bin/bash
# GEMM-only GPU check: smallest case first (a deadlock must not eat the budget), then the GEMM tests, then shapes bench.
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_ops_gpu.py -q -x -k "gemm_plain and 128-256-64" > gpurun_out/ops_first.log 2>&1
rc=$?
tail -15 gpurun_out/ops_first.log
if [ $rc -ne 0 ]; then echo "FIRST GEMM FAILED rc=$rc"; exit $rc; fi
timeout 600 python -m pytest tests/test_ops_gpu.py -q -x -k gemm > gpurun_out/ops_gemm.log 2>&1
rc=$?
echo "gemm tests rc=$rc"
tail -25 gpurun_out/ops_gemm.log
if [ $rc -ne 0 ]; then exit $rc; fi
timeout 600 python scripts/gemm_shapes_bench.py 0.25 > gpurun_out/gemm_shapes_bench.log 2>&1
echo "bench rc=$?"
cat gpurun_out/gemm_shapes_bench.log
