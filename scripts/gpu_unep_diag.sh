#!/bin/bash
# Where does the many-type (UNEP) step go?  Timing-only runs with parts of k_force_final switched off,
# then an `ncu --set full` capture of its six NEP kernels with source correlation.
set -u
T=${TAG:-r02_r}
mkdir -p gpurun_out
for c in "all:" "no_radial:B200MD_DEBUG_SKIP=1" "no_angular:B200MD_DEBUG_SKIP=2" "no_zbl:B200MD_DEBUG_SKIP=4"; do
  label=${c%%:*}; envs=${c#*:}
  env $envs timeout 300 python bench.py --workload unep --steps 20 --no-reference-gpu --no-cpu-baseline > gpurun_out/${T}_$label.json 2>/dev/null
  grep "^{" gpurun_out/${T}_$label.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$label', round(d['ms_per_step'],3), d['config'].get('nep_stage_ms'))"
done
timeout 600 ncu --set full --clock-control none --import-source on \
  -k regex:"k_force_final|k_desc_radial|k_split|k_mlp_tc|k_desc_angular|k_force_angular" -s 12 -c 6 \
  -o gpurun_out/${T}_prof_unep -f python bench.py --workload unep --steps 3 --warmup 3 --no-reference-gpu --no-cpu-baseline > gpurun_out/${T}_ncu.log 2>&1
tail -2 gpurun_out/${T}_ncu.log | cut -c1-200
ls -la gpurun_out/${T}_prof_unep.ncu-rep
