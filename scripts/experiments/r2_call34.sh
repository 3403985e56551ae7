#!/bin/bash
# tests + smoke + sanitizers on the final kernels, then the profile pass (summaries made on the box)
out=gpurun_out; mkdir -p $out
python __graft_entry__.py smoke 2>&1 | grep -v Using | tail -12 | tee $out/r2_s34_smoke.log
bash scripts/sanitize.sh r2_s34
BENCH=all KEEP_REPS="" bash scripts/gpu_profile.sh r2_s34 symik symik_f32 discrete continuous reachmap 2>&1 | tail -25
