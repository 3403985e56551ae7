#!/bin/bash
out=gpurun_out; mkdir -p $out
for v in t6 t7 t6c8 t6c7 t6; do
  echo "== $v" | tee -a $out/r2_s46_k3_occupancy.log
  python scripts/experiments/exp_r2_k3.py reachy2_symbolic_ik_b200/lib/variants/libr2ik_$v.so codes 2>&1 | grep "^codes" | tee -a $out/r2_s46_k3_occupancy.log
done
