#!/bin/bash
# A/B of the 128-bit coefficient loads (B200MD_NEP_CVEC) on the many-type (UNEP) and few-type (PbTe) paths,
# after the NEP parity tests.
set -u
T=${TAG:-r02_o}
mkdir -p gpurun_out
echo "== parity (NEP cases)"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference_properties.py -q -x -p no:cacheprovider -k "nep or golden or propert or invarian or consisten" > gpurun_out/${T}_pytest.txt 2>&1
tail -3 gpurun_out/${T}_pytest.txt
for v in 1 0; do
  echo "== unep CVEC=$v"
  B200MD_NEP_CVEC=$v timeout 600 python bench.py --workload unep --steps 30 --no-reference-gpu --no-cpu-baseline > gpurun_out/${T}_unep_cvec$v.json 2> gpurun_out/${T}_unep_cvec$v.err
  grep "^{" gpurun_out/${T}_unep_cvec$v.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('%.4g' % d['value'], d['ms_per_step'], d['config'].get('nep_stage_ms'))"
  tail -1 gpurun_out/${T}_unep_cvec$v.err | cut -c1-200
done
for v in 1 0; do
  echo "== pbte CVEC=$v"
  B200MD_NEP_CVEC=$v timeout 600 python bench.py --steps 60 --no-reference-gpu --no-cpu-baseline > gpurun_out/${T}_pbte_cvec$v.json 2> gpurun_out/${T}_pbte_cvec$v.err
  grep "^{" gpurun_out/${T}_pbte_cvec$v.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('%.4g' % d['value'], d['ms_per_step'], d['roofline']['stage_ms'])"
done
