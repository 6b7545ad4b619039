#!/bin/bash
# 2-GPU call (gpurun --gpus 2): reference NEP_MULTIGPU fixture, NCCL tests, bench at N = 1 and 2.
set -u
T=${TAG:-r02_two}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/${T}_gpus.txt 2>&1
echo "== reference gpumd, 2 GPUs (NEP_MULTIGPU) fixture"
timeout 900 python scripts/run_reference_gpumd.py --skip-speed --only 'md_pbte_(2gpu|lan|bao|npt)' 2>&1 | tail -40
for c in 2gpu lan bao npt; do
  cp gpurun_out/refgpu/md_pbte_$c/thermo.out gpurun_out/${T}_refgpu_md_pbte_${c}_thermo.out 2>/dev/null
  cp gpurun_out/refgpu/md_pbte_$c/thermo.out tests/golden/refgpu_md_pbte_${c}_thermo.out 2>/dev/null
done
echo "== pytest (multi-GPU tests)"
timeout 1500 python -m pytest tests/test_gpu_mgpu.py tests/test_gpu_domain.py tests/test_gpu_stochastic.py tests/test_gpu_md.py -m gpu -q 2>&1 | tail -45 | tee gpurun_out/${T}_pytest.txt
for g in ${BENCH_GPUS:-1 2}; do
  echo "== bench N=$g"
  if [ "$g" = "1" ]; then
    timeout 900 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err
  else
    timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 2951$g \
      bench.py --gpus $g --steps 100 --warmup 5 ${BENCH_ARGS:-} > gpurun_out/${T}_bench_n$g.json 2> gpurun_out/${T}_bench_n$g.err
  fi
  python - <<PY
import json
for ln in open("gpurun_out/${T}_bench_n$g.json"):
    if ln.startswith("{"):
        d=json.loads(ln); print("N=$g value %.4g ms/step %.3f e2e %.4g launches %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"]))
        print(d["config"].get("phase_ms_rank0"), d["config"].get("max_owned"), d["config"].get("max_local_with_ghosts"))
        r=d.get("reference_gpu"); print("reference_gpu", r and (r.get("value"), r.get("path"), r.get("wall_s")))
PY
  tail -4 gpurun_out/${T}_bench_n$g.err
done
ls gpurun_out | tail -30
