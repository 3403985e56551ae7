#!/bin/bash
# On the GPU box: rebuild libr2ik.so with each K1 occupancy variant and time the symik workload.
out=gpurun_out; mkdir -p $out
for mb in ${@:-3 4 5 6}; do
  R2IK_NVCC_EXTRA="-DR2IK_K1_MINBLOCKS=$mb" python -m reachy2_symbolic_ik_b200.build --force > /dev/null
  echo -n "minblocks=$mb " ; python bench.py --steps 300 --warmup 10 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value %.4e kernel_ms %.4f fp64frac %.3f parity_over %s maxerr %.2e'%(d['value'],d['roofline']['kernel_ms'],d['roofline_fp64']['frac'],d['parity']['over_1e-9'],d['parity']['max_abs_err_joints_rad']))"
done | tee $out/tune_k1.txt
python -m reachy2_symbolic_ik_b200.build --force > /dev/null
