#!/bin/bash
# Secondary workloads (C2 LJ, C4 UNEP share, C5 Si share) on 1 GPU: bench line, ncu launch list and one
# full capture of the dominant kernels each.   gpurun --timeout 2400 -- bash scripts/gpu_side.sh
set -u
T=${TAG:-r02_side}
mkdir -p gpurun_out
for w in ${WORKLOADS:-unep lj si}; do
  echo "== bench $w"
  timeout 600 python bench.py --workload $w --steps ${SIDE_STEPS:-50} --warmup 5 > gpurun_out/${T}_bench_$w.json 2> gpurun_out/${T}_bench_$w.err
  cut -c1-900 gpurun_out/${T}_bench_$w.json; tail -3 gpurun_out/${T}_bench_$w.err
  echo "== ncu launch list $w"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/${T}_launches_$w.csv \
    python bench.py --workload $w --steps 3 --warmup 3 > gpurun_out/${T}_ncu_$w.log 2>&1
  tail -2 gpurun_out/${T}_ncu_$w.log | cut -c1-200
  case $w in
    unep) K='k_force_final|k_force_angular|k_desc_radial|k_desc_angular|k_mlp_tc|k_split' ;;
    lj) K='k_lj|k_skin_list' ;;
    si) K='k_tersoff' ;;
  esac
  echo "== ncu full $w ($K)"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$K" -s ${NCU_SKIP:-8} -c ${NCU_COUNT:-6} \
    -o gpurun_out/${T}_prof_$w -f python bench.py --workload $w --steps 3 --warmup 3 > gpurun_out/${T}_ncufull_$w.log 2>&1
  tail -2 gpurun_out/${T}_ncufull_$w.log | cut -c1-200
done
ls -la gpurun_out | tail -20
