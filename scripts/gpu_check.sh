#!/bin/bash
# One gpurun call: smoke, GPU parity tests, a short bench, an ncu launch list and full captures of
# the dominant kernels.  Everything lands in gpurun_out/.
#   /usr/local/graft/bin/gpurun --timeout 1700 -- bash scripts/gpu_check.sh
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q ${PYTEST_ARGS:-} 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.txt
echo "== bench"; timeout 600 python bench.py --steps ${BENCH_STEPS:-100} --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
for w in ${SIDE_WORKLOADS:-lj}; do
echo "== bench $w"; timeout 400 python bench.py --workload $w --steps ${SIDE_STEPS:-100} --warmup 5 > gpurun_out/bench_$w.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_$w.json
done
if [ "${RUN_REF:-0}" = "1" ]; then
echo "== reference gpumd"; timeout 1500 python scripts/run_reference_gpumd.py ${REF_ARGS:-} 2>&1 | tail -60
fi
if [ "${SKIP_NCU:-0}" != "1" ]; then
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -3 gpurun_out/ncu_bench.log
echo "== ncu full: radial kernels"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"${NCU_KERNELS:-k_force_final|k_desc_radial|k_mlp_tc}" -s ${NCU_SKIP:-6} -c ${NCU_COUNT:-6} \
  -o gpurun_out/prof_radial -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
fi
ls -la gpurun_out
