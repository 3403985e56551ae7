#!/bin/bash
# Run on the GPU box (gpurun): tests, the four bench workloads, ncu launch lists and one --set full capture per kernel.
# usage: scripts/gpu_profile.sh <tag> [workloads...]
tag=${1:-r1}; shift
wls=${@:-symik discrete continuous reachmap}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -q 2>&1 | tail -15 > $out/${tag}_pytest.log
tail -3 $out/${tag}_pytest.log
declare -A KRE=( [symik]=k_symik_solve [symik_f32]=k_symik_.*_f32 [discrete]=k_ctl_discrete [continuous]=k_cont_ [reachmap]=k_reach_map )
declare -A NCAP=( [continuous]=5 [symik_f32]=2 )
python -c "import bench; print(bench.csrc_sha16())" > $out/${tag}_csrc_sha16.txt 2>/dev/null
for w in $wls; do
  python bench.py --workload $w > $out/${tag}_bench_${w}.json 2> $out/${tag}_bench_${w}.err
  cut -c1-400 $out/${tag}_bench_${w}.json
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_${w}.csv \
      python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:${KRE[$w]} -s 4 -c ${NCAP[$w]:-1} -f -o $out/${tag}_${w} \
      python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
done
ls -la $out | tail -20
