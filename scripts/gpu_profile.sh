#!/bin/bash
# Run on the GPU box (gpurun): tests, the bench workloads, ncu launch lists and one --set full capture per kernel.
# The captures are summarised ON THE BOX (gpurun brings back at most 64 MiB): <tag>_<wl>_ncu_full.txt, <tag>_<wl>_ncu.json;
# only the .ncu-rep files named in KEEP_REPS travel back.
# usage: [BENCH=per|all|none] [KEEP_REPS="continuous"] scripts/gpu_profile.sh <tag> [workloads...]
tag=${1:-r1}; shift
wls=${@:-symik symik_f32 discrete continuous reachmap}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -q 2>&1 | tail -15 > $out/${tag}_pytest.log
tail -3 $out/${tag}_pytest.log
declare -A KRE=( [symik]=k_symik_solve [symik_f32]=k_symik_.*_f32 [discrete]=k_disc_ [continuous]=k_cont_ [reachmap]=k_reach_map )
declare -A KNAME=( [symik]=k_symik_solve [symik_f32]=k_symik_ [discrete]=k_disc_ [continuous]=k_cont_ [reachmap]=k_reach_map )
declare -A NCAP=( [continuous]=5 [symik_f32]=2 [discrete]=4 )
declare -A SUM=( [continuous]=sum [discrete]=sum [symik_f32]=sum )
python -c "import bench; print(bench.csrc_sha16())" > $out/${tag}_csrc_sha16.txt 2>/dev/null
if [ "${BENCH:-per}" = all ]; then
  python bench.py --steps 20 --warmup 5 > $out/${tag}_bench_all.json 2> $out/${tag}_bench_all.err
  cut -c1-300 $out/${tag}_bench_all.json
fi
for w in $wls; do
  if [ "${BENCH:-per}" = per ]; then
    python bench.py --workload $w > $out/${tag}_bench_${w}.json 2> $out/${tag}_bench_${w}.err
    cut -c1-400 $out/${tag}_bench_${w}.json
  fi
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_${w}.csv \
      python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:${KRE[$w]} -s 4 -c ${NCAP[$w]:-1} -f -o $out/${tag}_${w} \
      python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
  python scripts/ncu_summary.py $out/${tag}_${w}.ncu-rep $out/${tag}_${w}_ncu_full.txt > /dev/null 2>&1
  python scripts/ncu_to_json.py $out/${tag}_${w}.ncu-rep $w ${KNAME[$w]} $tag ${SUM[$w]} > /dev/null 2>&1 && cp profiles/${w}_ncu.json $out/${tag}_${w}_ncu.json
  case " ${KEEP_REPS:-} " in *" $w "*) ;; *) rm -f $out/${tag}_${w}.ncu-rep ;; esac
done
du -sh $out; ls -la $out | tail -30
