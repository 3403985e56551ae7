#!/bin/bash
out=gpurun_out; mkdir -p $out
python -m pytest tests/test_gpu_api_r2.py -m gpu -q -x -k "legacy" 2>&1 | tail -2
for v in base b64m8 b64m10 b256m2 m5 m6 b32m16; do
  python scripts/experiments/exp_r2_k1.py reachy2_symbolic_ik_b200/lib/variants/libr2ik_$v.so 2>&1 | grep "mat4 \|euler6  outputs lean" | tee -a $out/r2_s37_k1_shapes.log
done
