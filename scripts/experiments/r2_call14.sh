mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/r2_s14_pytest_2gpu.log
python scripts/experiments/exp_r2_e2e_nrank.py 2>/dev/null | tail -1 | tee gpurun_out/r2_s14_e2e_1rank.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 scripts/experiments/exp_r2_e2e_nrank.py 2>/dev/null | tail -1 | tee gpurun_out/r2_s14_e2e_2rank.json
nproc; lscpu | grep -E "Model name|Socket|NUMA node\(s\)|^CPU\(s\)"
