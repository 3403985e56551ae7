mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/r2_s26_pytest.log
python scripts/experiments/exp_r2_k3.py codes phased4 2>&1 | grep -v "^Using" | tee gpurun_out/r2_s26_k3.log
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_cont -s 15 -c 5 python scripts/experiments/exp_r2_k3.py codes 2>&1 | grep -E "k_cont_[a-z_]*\(|gpu__time|dram__|smsp__" | sed 's/(ArmConst.*//; s/(long.*//' | paste - - - - - - | tee -a gpurun_out/r2_s26_k3.log
