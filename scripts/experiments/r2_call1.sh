set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2_s1_pytest.log; tail -5 gpurun_out/r2_s1_pytest.log
python scripts/experiments/exp_r2_scalar_latency.py 300 > gpurun_out/r2_s1_scalar_latency.json 2> gpurun_out/r2_s1_scalar_latency.err; cat gpurun_out/r2_s1_scalar_latency.json; tail -3 gpurun_out/r2_s1_scalar_latency.err
for g in 0 32 128; do python scripts/experiments/exp_r2_l2fetch.py $g 2>&1 | tail -2; done | tee gpurun_out/r2_s1_l2fetch.log
for g in 0 32; do ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_symik_solve -s 3 -c 1 python scripts/experiments/exp_r2_l2fetch.py $g 2>&1 | grep -E "dram__|gpu__time|granularity" ; done | tee gpurun_out/r2_s1_l2fetch_ncu.log
nvidia-smi topo -m > gpurun_out/r2_s1_topo.txt 2>&1; nproc; free -g | head -2
