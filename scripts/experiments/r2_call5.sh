mkdir -p gpurun_out
python -m pytest tests/test_gpu_control.py tests/test_gpu_zz_overrides.py -q -x 2>&1 | tail -5
python scripts/experiments/exp_r2_k3.py tiled phased4 2>&1 | grep -v "^Using" | tee gpurun_out/r2_s5_k3.log
for v in tile3 tile5; do python scripts/experiments/exp_r2_k3.py reachy2_symbolic_ik_b200/lib/variants/libr2ik_$v.so tiled 2>&1 | grep -v "^Using" | sed "s/^/$v /" | tee -a gpurun_out/r2_s5_k3.log; done
ncu --set full --clock-control none --import-source on -k regex:k_cont_joints_finish -s 2 -c 1 -f -o gpurun_out/r2_s5_k3_jf python scripts/experiments/exp_r2_k3.py tiled > /dev/null 2>&1
ncu -i gpurun_out/r2_s5_k3_jf.ncu-rep --page raw --csv | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); hdr=rows[0]; v=rows[2]
for k in ['gpu__time_duration.sum','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','dram__bytes_read.sum','dram__bytes_write.sum','sm__warps_active.avg.pct_of_peak_sustained_active']:
    print(k, v[hdr.index(k)], rows[1][hdr.index(k)])
for i,h in enumerate(hdr):
    if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio') and float(v[i])>0.1: print(h[34:-23], v[i])
" | tee -a gpurun_out/r2_s5_k3.log
