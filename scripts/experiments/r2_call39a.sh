#!/bin/bash
out=gpurun_out; mkdir -p $out
python __graft_entry__.py smoke 2>&1 | grep -v Using | tail -12 | tee $out/r2_s48_smoke.log
bash scripts/sanitize.sh r2_s48
