#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_disc_(classify|search|finish)" -f -o $out/r2_s31_k2c python scripts/experiments/exp_r2_k2.py --once > /dev/null 2>&1
python scripts/ncu_summary.py $out/r2_s31_k2c.ncu-rep $out/r2_s31_k2c_ncu_full.txt > /dev/null
cat $out/r2_s31_k2c_ncu_full.txt
