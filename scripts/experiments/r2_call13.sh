mkdir -p gpurun_out
nvidia-smi -L | head -3
python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/r2_s13_pytest_2gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_s13_bench_2gpu.json 2> gpurun_out/r2_s13_bench_2gpu.err
tail -3 gpurun_out/r2_s13_bench_2gpu.err; cut -c1-300 gpurun_out/r2_s13_bench_2gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 scripts/experiments/exp_pcie_nrank.py > gpurun_out/r2_s13_pcie_2rank.json 2>/dev/null; cat gpurun_out/r2_s13_pcie_2rank.json
