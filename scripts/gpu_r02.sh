#!/bin/bash
# Round-2 GPU check: parity tests, A/B bench of the radial kernels, diagnostics, ncu captures.
#   /usr/local/graft/bin/gpurun --timeout 2400 -- bash scripts/gpu_r02.sh
set -u
T=${TAG:-r02}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest -m gpu ${PYTEST_TARGET:-tests}"
timeout ${PYTEST_TIMEOUT:-1200} python -m pytest ${PYTEST_TARGET:-tests} -m gpu -q ${PYTEST_ARGS:-} 2>&1 | tail -${PYTEST_TAIL:-40} | tee gpurun_out/${T}_pytest_gpu.txt
# BENCH_ENVS: space-separated "label:VAR=val[,VAR=val]" items, one bench run each
for item in ${BENCH_ENVS:-default:B200MD_NEP_RADIAL=v2}; do
  v=${item%%:*}; envs=${item#*:}
  echo "== bench $v ($envs)"
  env $(echo $envs | tr ',' ' ') timeout 900 python bench.py --steps ${BENCH_STEPS:-100} --warmup 5 ${BENCH_ARGS:-} > gpurun_out/${T}_bench_$v.json 2> gpurun_out/${T}_bench_$v.err
  python - <<PY
import json
for ln in open("gpurun_out/${T}_bench_$v.json"):
    if ln.startswith("{"):
        d=json.loads(ln); print("$v value %.4g ms/step %.3f e2e %.4g" % (d["value"], d["ms_per_step"], d["e2e"]["value"]))
        if d.get("roofline"): print(d["roofline"].get("stage_ms"))
        if d.get("neighbor"): print(d["neighbor"])
        if d.get("reference_gpu"): print("reference_gpu", d["reference_gpu"].get("value"), d["reference_gpu"].get("wall_s"))
PY
  tail -3 gpurun_out/${T}_bench_$v.err
done
if [ -n "${DIAG:-}" ]; then
  echo "== diag $DIAG"; timeout 900 python scripts/diag_r02.py $DIAG 2>&1 | tail -30 | cut -c1-400
fi
if [ "${SKIP_NCU:-0}" != "1" ]; then
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_launches.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline ${BENCH_ARGS:-} > gpurun_out/${T}_ncu_bench.log 2>&1
tail -3 gpurun_out/${T}_ncu_bench.log
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"${NCU_KERNELS:-k_force_final|k_desc_radial|k_mlp_tc}" -s ${NCU_SKIP:-6} -c ${NCU_COUNT:-6} \
  -o gpurun_out/${T}_prof -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline ${BENCH_ARGS:-} > gpurun_out/${T}_ncu_full.log 2>&1
tail -3 gpurun_out/${T}_ncu_full.log
fi
ls -la gpurun_out | tail -20
