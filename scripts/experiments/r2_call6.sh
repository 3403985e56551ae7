mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -5
python scripts/experiments/exp_r2_e2e.py 2>&1 | grep -v "^Using" | tee gpurun_out/r2_s6_e2e.log
