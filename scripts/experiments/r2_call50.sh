#!/bin/bash
# 8 GPUs: the default bench on the final sources
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_s50_bench_all_8gpu.json 2> gpurun_out/r2_s50_bench_8gpu.err
tail -1 gpurun_out/r2_s50_bench_8gpu.err; cut -c1-200 gpurun_out/r2_s50_bench_all_8gpu.json
