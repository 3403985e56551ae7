mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_s17_topo_8gpu.txt 2>&1; nproc; free -g | sed -n 2p
for n in 8 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2962$n scripts/experiments/exp_pcie_nrank.py 2>/dev/null | tail -1 > gpurun_out/r2_s17_pcie_${n}rank.json; cat gpurun_out/r2_s17_pcie_${n}rank.json | cut -c1-420
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2963$n scripts/experiments/exp_r2_e2e_nrank.py 2>/dev/null | tail -1 > gpurun_out/r2_s17_e2e_${n}rank.json; cat gpurun_out/r2_s17_e2e_${n}rank.json
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_s17_bench_8gpu.json 2> gpurun_out/r2_s17_bench_8gpu.err
tail -2 gpurun_out/r2_s17_bench_8gpu.err; cut -c1-200 gpurun_out/r2_s17_bench_8gpu.json
python -m pytest tests/test_gpu_api_r2.py -m gpu -q -k "multi_gpu" 2>&1 | tail -2
