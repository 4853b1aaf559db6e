#!/bin/bash
# one ncu --set full capture of the tcgen05 attention kernel per shape (after the plain run exited 0)
mkdir -p gpurun_out
for shp in dino qwen; do
  timeout 120 python scripts/attn_bench.py $shp > gpurun_out/attn_bench_$shp.log 2>&1 || { echo "plain run failed"; cat gpurun_out/attn_bench_$shp.log; exit 1; }
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:fa_tcgen05 -s 3 -c 1 -o gpurun_out/attn_$shp -f python scripts/attn_bench.py $shp > gpurun_out/attn_ncu_$shp.log 2>&1
  echo "ncu $shp rc=$?"; tail -3 gpurun_out/attn_ncu_$shp.log
done
ls -la gpurun_out/*.ncu-rep
