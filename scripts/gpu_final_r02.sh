#!/bin/bash
# End-of-round GPU call on 1x B200: full GPU test suite, smoke(), the headline bench line (with the
# reference gpumd beside it), the UNEP stage breakdown, ncu launch list + full capture of the five NEP
# kernels, and an ncu capture of the EAM and neighbour-rebuild kernels.
set -u
T=${TAG:-r02_p}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${T}_gpu.txt 2>&1
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests/ -q -m gpu -p no:cacheprovider > gpurun_out/${T}_pytest_gpu.txt 2>&1
tail -5 gpurun_out/${T}_pytest_gpu.txt
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
echo "== bench (headline)"
timeout 900 python bench.py > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err
python - <<PY
import json
for ln in open("gpurun_out/${T}_bench_n1.json"):
    if ln.startswith("{"):
        d = json.loads(ln)
        print("value %.4g ms/step %.3f e2e %.4g" % (d["value"], d["ms_per_step"], d["e2e"]["value"]))
        print(d["roofline"]["stage_ms"]); print("reference_gpu", (d.get("reference_gpu") or {}).get("value"))
        print("cpu_baseline", d.get("cpu_baseline", {}).get("value"), "neighbor", d.get("neighbor", {}).get("rebuild_ms"))
PY
tail -2 gpurun_out/${T}_bench_n1.err | cut -c1-300
echo "== bench --impl reference"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_ref.json 2>/dev/null; cut -c1-300 gpurun_out/${T}_bench_ref.json
echo "== bench unep"
timeout 600 python bench.py --workload unep --steps 30 --no-reference-gpu > gpurun_out/${T}_bench_unep.json 2> gpurun_out/${T}_bench_unep.err
grep "^{" gpurun_out/${T}_bench_unep.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['config'].get('nep_stage_ms'))"
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_launches.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/${T}_ncu_bench.log 2>&1
tail -2 gpurun_out/${T}_ncu_bench.log | cut -c1-200
echo "== ncu full (NEP kernels)"
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:"k_force_final|k_desc_radial|k_mlp_tc|k_desc_angular|k_force_angular" -s 10 -c 10 \
  -o gpurun_out/${T}_prof -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/${T}_ncu_full.log 2>&1
tail -2 gpurun_out/${T}_ncu_full.log | cut -c1-200
echo "== EAM + rebuild"
timeout 300 python scripts/prof_eam_rebuild.py 2>&1 | tail -3
timeout 600 ncu --set full --clock-control none --import-source on \
  -k regex:"k_eam|k_skin_list|k_cell|k_sort_cells|k_commit|k_scan_cells|k_pack_check" -c 12 \
  -o gpurun_out/${T}_prof_eam_rebuild -f python scripts/prof_eam_rebuild.py > gpurun_out/${T}_ncu_eam.log 2>&1
tail -2 gpurun_out/${T}_ncu_eam.log | cut -c1-200
ls -la gpurun_out | grep ${T}
