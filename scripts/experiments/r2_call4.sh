mkdir -p gpurun_out
python -m pytest tests/test_gpu_control.py tests/test_gpu_zz_overrides.py -q -x 2>&1 | tail -5
python scripts/experiments/exp_r2_k3.py tiled phased4 2>&1 | grep -v "^Using" | tee gpurun_out/r2_s4_k3.log
for v in tile3 tile5; do python scripts/experiments/exp_r2_k3.py reachy2_symbolic_ik_b200/lib/variants/libr2ik_$v.so tiled 2>&1 | grep -v "^Using" | sed "s/^/$v /" | tee -a gpurun_out/r2_s4_k3.log; done
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_cont -s 8 -c 6 python scripts/experiments/exp_r2_k3.py tiled 2>&1 | grep -E "k_cont|gpu__time" | paste - - | tee -a gpurun_out/r2_s4_k3.log
ncu --set full --clock-control none --import-source on -k regex:k_cont_joints_finish -s 2 -c 1 -f -o gpurun_out/r2_s4_k3_jf python scripts/experiments/exp_r2_k3.py tiled > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
