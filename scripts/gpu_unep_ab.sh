#!/bin/bash
# A/B of the many-type (UNEP) tuning switches, after the NEP parity tests.  Usage:
#   TAG=r02_q CASES="label:ENV=val,ENV2=val label2:" bash scripts/gpu_unep_ab.sh
set -u
T=${TAG:-r02_o}
CASES=${CASES:-"default: radreg0:B200MD_NEP_RADREG=0 cstage0:B200MD_NEP_CSTAGE=0 ff_depth2:B200MD_NEP_VARIANT=2 ff_6blocks:B200MD_NEP_VARIANT=1"}
PB_CASES=${PB_CASES:-"default: cstage0:B200MD_NEP_CSTAGE=0"}
mkdir -p gpurun_out
echo "== parity (NEP cases)"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference_properties.py -q -x -p no:cacheprovider -k "nep or golden or propert or invarian or consisten" > gpurun_out/${T}_pytest.txt 2>&1
tail -3 gpurun_out/${T}_pytest.txt
timeout 600 python -m pytest tests/test_gpu_mgpu.py -q -x -p no:cacheprovider -k "many_types or single_point" > gpurun_out/${T}_pytest_mgpu.txt 2>&1
tail -3 gpurun_out/${T}_pytest_mgpu.txt
run() { # workload label envs extra-args
  local w=$1 c=$2; shift 2
  local label=${c%%:*} envs=${c#*:}
  echo "== $w $label ($envs)"
  env $(echo $envs | tr ',' ' ') timeout 600 python bench.py "$@" --no-reference-gpu --no-cpu-baseline > gpurun_out/${T}_${w}_$label.json 2> gpurun_out/${T}_${w}_$label.err
  grep "^{" gpurun_out/${T}_${w}_$label.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('%.4g' % d['value'], round(d['ms_per_step'],4), d['config'].get('nep_stage_ms') or d['roofline']['stage_ms'])"
  tail -1 gpurun_out/${T}_${w}_$label.err | cut -c1-200
}
for c in $CASES; do run unep $c --workload unep --steps 30; done
for c in $PB_CASES; do run pbte $c --steps 60; done
