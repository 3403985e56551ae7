mkdir -p gpurun_out
python -m pytest tests/test_gpu_control.py tests/test_gpu_zz_overrides.py -q -x 2>&1 | tail -15
python scripts/experiments/exp_r2_k3.py 2>&1 | grep -v "^Using" | tee gpurun_out/r2_s3_k3.log
