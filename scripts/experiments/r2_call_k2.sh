#!/bin/bash
# K2 compacted passes with programmatic dependent launch
out=gpurun_out; mkdir -p $out
python -m pytest tests/test_gpu_api_r2.py tests/test_gpu_control.py -m gpu -q -x -k "discrete" 2>&1 | tail -3
python scripts/experiments/exp_r2_k2.py 2>&1 | grep -v Using | tee $out/r2_s33_k2.log
python scripts/experiments/exp_r2_k2.py reachy2_symbolic_ik_b200/lib/variants/libr2ik_k2c_nopdl.so 2>&1 | grep "compact=True\|identical\|n = " | tee -a $out/r2_s33_k2.log
compute-sanitizer --tool racecheck python scripts/experiments/exp_r2_k2.py --once 2>&1 | tail -4
compute-sanitizer --tool memcheck python scripts/experiments/exp_r2_k2.py --once 2>&1 | tail -4
