#!/bin/bash
out=gpurun_out; mkdir -p $out
python -m pytest tests/test_gpu_api_r2.py tests/test_gpu_control.py tests/test_gpu_symik.py -m gpu -q -x -k "discrete or fp32 or f32" 2>&1 | tail -3
for v in k2c_fullsolve k2c_circle k2c_fullsolve k2c_circle; do
  python scripts/experiments/exp_r2_k2.py reachy2_symbolic_ik_b200/lib/variants/libr2ik_$v.so 2>&1 | grep "compact=True\|identical" | tee -a $out/r2_s40_k2.log
done
python scripts/experiments/bench_with_lib.py reachy2_symbolic_ik_b200/lib/libr2ik.so --workload symik_f32 --steps 2000 --warmup 20 --no-cpu-baseline 2>/dev/null | cut -c1-200 | tee -a $out/r2_s40_k2.log
