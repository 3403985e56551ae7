mkdir -p gpurun_out
python scripts/experiments/exp_r2_k3.py phased4 codes 2>&1 | grep -v "^Using" | tee gpurun_out/r2_s21_k3.log
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_cont -s 10 -c 4 python scripts/experiments/exp_r2_k3.py codes 2>&1 | grep -E "k_cont_[a-z_]*\(|gpu__time|dram__" | sed 's/(ArmConst.*//' | paste - - - - | tee -a gpurun_out/r2_s21_k3.log
