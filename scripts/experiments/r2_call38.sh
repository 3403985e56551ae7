#!/bin/bash
out=gpurun_out; mkdir -p $out
for v in base k1pdl base k1pdl; do
  python scripts/experiments/bench_with_lib.py reachy2_symbolic_ik_b200/lib/variants/libr2ik_$v.so --workload symik --steps 2000 --warmup 20 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', '%.4g'%d['value'], '%.4f ms'%d['ms_per_step'], 'sustained %.4g'%d.get('sustained',{}).get('value',0), 'e2e %.3g'%d['e2e']['value'], 'lean %.3g'%d['e2e']['lean']['value'], d['parity']['state_mismatches'])" | tee -a $out/r2_s38_k1_pdl.log
done
