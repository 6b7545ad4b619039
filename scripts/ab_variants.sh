#!/bin/bash
# A/B of kernel tuning variants (B200MD_NEP_VARIANT / B200MD_NEP_TEAM) on the headline workload
CONFIGS=${AB_CONFIGS:-0:0 0:2}
for cfg in $CONFIGS; do
  T=${cfg%%:*}; V=${cfg##*:}
  echo "TEAM=$T VARIANT=$V"
  B200MD_NEP_TEAM=$T B200MD_NEP_VARIANT=$V timeout 300 python bench.py --steps 60 --warmup 5 --no-cpu-baseline | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.4g'%d['value'], round(d['ms_per_step'],3), d['roofline']['stage_ms'])"
done
