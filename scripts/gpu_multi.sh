#!/bin/bash
# Multi-GPU check (gpurun --gpus N): domain-decomposition test + bench at 1..N GPUs.
set -u
N=${NGPUS:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv > gpurun_out/gpus.txt 2>&1
echo "== pytest gpu (all)"; timeout 1200 python -m pytest ${PYTEST_TARGET:-tests} -m gpu -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_multi.txt
for g in ${BENCH_GPUS:-1 2}; do
  echo "== bench N=$g"
  if [ "$g" = "1" ]; then
    timeout 600 python bench.py --steps 100 --warmup 5 > gpurun_out/bench_n1.json 2>> gpurun_out/bench_multi.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 2951$g \
      bench.py --gpus $g --steps 100 --warmup 5 > gpurun_out/bench_n$g.json 2>> gpurun_out/bench_multi.err
  fi
  python - <<PY
import json
for ln in open("gpurun_out/bench_n$g.json"):
    if ln.startswith("{"):
        d=json.loads(ln); print("N=$g value %.4g ms/step %.3f e2e %.4g launches %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"]))
        if d.get("roofline"): print(d["roofline"]["stage_ms"])
PY
done
tail -5 gpurun_out/bench_multi.err
