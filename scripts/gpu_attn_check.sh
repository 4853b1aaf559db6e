#!/bin/bash
# tcgen05 attention check: parity tests under a timeout (a deadlock must not eat the budget), then timings.
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_ops_gpu.py -q -x -k "attention" > gpurun_out/attn_all.log 2>&1
rc=$?
echo "attn_all rc=$rc"
tail -30 gpurun_out/attn_all.log
if [ $rc -ne 0 ]; then exit $rc; fi
for d in 0 1 7; do echo "== VLA_FA_DEBUG=$d"; VLA_FA_DEBUG=$d timeout 120 python scripts/attn_bench.py 2>&1 | sed 's/mma.sync.*| tcgen05/tcgen05/'; done | tee gpurun_out/attn_dbg.log
