#!/bin/bash
# full GPU suite + smoke + headline bench (used after a late kernel change)
set -u
T=${TAG:-r02_t}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -q -m gpu -p no:cacheprovider > gpurun_out/${T}_pytest_gpu.txt 2>&1
tail -3 gpurun_out/${T}_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err
python - <<PY
import json
for ln in open("gpurun_out/${T}_bench_n1.json"):
    if ln.startswith("{"):
        d = json.loads(ln)
        print("value %.4g ms/step %.3f e2e %.4g" % (d["value"], d["ms_per_step"], d["e2e"]["value"]))
        print(d["roofline"]["stage_ms"]); print("reference_gpu", (d.get("reference_gpu") or {}).get("value"))
PY
timeout 300 python bench.py --workload unep --steps 30 --no-reference-gpu --no-cpu-baseline 2>/dev/null | grep "^{" > gpurun_out/${T}_bench_unep.json
python -c "
import json; d=json.load(open('gpurun_out/${T}_bench_unep.json')); print(d['value'], d['ms_per_step'], d['config']['nep_stage_ms'])"
