mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2_s2_pytest.log; tail -5 gpurun_out/r2_s2_pytest.log
( time python bench.py --steps 20 --warmup 5 > gpurun_out/r2_s2_bench_all.json 2> gpurun_out/r2_s2_bench_all.err ) 2>&1 | tail -3
tail -5 gpurun_out/r2_s2_bench_all.err; cut -c1-600 gpurun_out/r2_s2_bench_all.json
( time python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_s2_bench_ref.json 2> gpurun_out/r2_s2_bench_ref.err ) 2>&1 | tail -3
cut -c1-300 gpurun_out/r2_s2_bench_ref.json
python scripts/experiments/exp_pcie_nrank.py > gpurun_out/r2_s2_pcie_1rank.json 2>&1; cat gpurun_out/r2_s2_pcie_1rank.json
