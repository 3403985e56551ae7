mkdir -p gpurun_out
python -m pytest tests/test_gpu_symik.py tests/test_gpu_api_r2.py -m gpu -q -x 2>&1 | tail -3
python scripts/experiments/exp_r2_k1.py 2>&1 | grep -v "^Using" | tee gpurun_out/r2_s12_k1.log
python scripts/experiments/exp_r2_k1.py reachy2_symbolic_ik_b200/lib/variants/libr2ik_staged.so 2>&1 | grep -v "^Using" | tee -a gpurun_out/r2_s12_k1.log
M="gpu__time_duration.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_global_op_st.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum,lts__t_sectors_op_write.sum"
for lib in "" reachy2_symbolic_ik_b200/lib/variants/libr2ik_staged.so; do
  ncu --metrics $M --clock-control none -k regex:k_symik_solve -s 3 -c 1 --csv python scripts/experiments/exp_r2_k1.py $lib 2>/dev/null | grep -E "k_symik_solve" | awk -F'","' '{print $5" | "$(NF-2)" | "$(NF)}' | sed "s/^/[${lib##*_}] /" | tee -a gpurun_out/r2_s12_k1.log
done
