#!/bin/bash
# 8 GPUs: the default bench (every workload; the reach map reports its two sharding forms) on the final kernels
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_s35_bench_all_8gpu.json 2> gpurun_out/r2_s35_bench_8gpu.err
tail -2 gpurun_out/r2_s35_bench_8gpu.err; cut -c1-300 gpurun_out/r2_s35_bench_all_8gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29642 bench.py --gpus 4 --steps 20 --warmup 5 --workload reachmap > gpurun_out/r2_s35_bench_reachmap_4gpu.json 2>/dev/null
cut -c1-200 gpurun_out/r2_s35_bench_reachmap_4gpu.json
