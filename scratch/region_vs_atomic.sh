#!/bin/bash
# K3s (shared-memory regions) against the L2-atomic K3 on config 2: bench lines without the stage extras
for rb in 12 0; do
  PG_REGION_BITS=$rb python bench.py --steps 20 --no-cpu-baseline --no-stages > gpurun_out/r2f_bench_rb$rb.json 2> gpurun_out/r2f_bench_rb$rb.err
  echo "rb=$rb rc=$?"
done
