#!/bin/bash
# usage: scripts/exp_bench_variants.sh <workload> : bench.py --workload W on the in-tree lib and every lib/variants/*.so (development aid)
w=$1
for lib in "" reachy2_symbolic_ik_b200/lib/variants/*.so; do
  [ -n "$lib" ] && export R2IK_LIB=$PWD/$lib
  timeout 300 python bench.py --workload $w --no-cpu-baseline 2>/dev/null | python -c "
import json,sys,os
d=json.loads(sys.stdin.read()); print(os.path.basename(os.environ.get('R2IK_LIB','in-tree')), '%.4g'%d['value'], '%.3f ms'%d['ms_per_step'], 'e2e %.3g'%d['e2e']['value'], (d.get('parity') or {}).get('max_abs_err_joints_rad'), (d.get('parity') or {}).get('state_mismatches'))"
done
