#!/bin/bash
# final kernels: tests + smoke + sanitizers, then the profile pass (summaries made on the box)
out=gpurun_out; mkdir -p $out
python __graft_entry__.py smoke 2>&1 | grep -v Using | tail -12 | tee $out/r2_s39_smoke.log
bash scripts/sanitize.sh r2_s39
BENCH=all KEEP_REPS="" bash scripts/gpu_profile.sh r2_s39 symik symik_f32 discrete continuous reachmap 2>&1 | tail -12
