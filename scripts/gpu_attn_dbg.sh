#!/bin/bash
mkdir -p gpurun_out
for d in 0 1 2 4 3 6 7; do echo "== VLA_FA_DEBUG=$d"; VLA_FA_DEBUG=$d timeout 120 python scripts/attn_bench.py 2>&1 | sed 's/mma.sync.*| tcgen05/tcgen05/'; done | tee gpurun_out/attn_dbg.log
