mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/r2_s27_pytest.log
bash scripts/sanitize.sh r2_s27
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_s27_bench_all.json 2> gpurun_out/r2_s27_bench_all.err; tail -2 gpurun_out/r2_s27_bench_all.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2_s27_bench_all.json"))
print("value %.4g e2e %.4g" % (d["value"], d["e2e"]["value"]), {k:"%.4g"%v.get("value",0) for k,v in d["e2e"].items() if isinstance(v,dict)})
for k,v in d["workloads"].items():
    print(k, v.get("error") or ("%.4g"%v["value"], "%.4g ms"%v["ms_per_step"], "e2e %.4g"%v["e2e"]["value"], v.get("parity")))
PY
