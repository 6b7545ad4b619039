#!/bin/bash
# Multi-GPU call (gpurun --gpus N): BASELINE configs at their stated GPU counts + scaling lines.
#   N=8: C3 weak (slabs) with the reference's NEP_MULTIGPU beside it, C4 (8 M UNEP nvt_ber), C3 strong (1 M, 2x2x2)
#   N=4: C5 (2 M Si Tersoff + heat current), C3 weak
set -u
N=${NGPUS:-8}
T=${TAG:-r02_m}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/${T}_gpus.txt 2>&1
run() { # label, extra args
  local label=$1; shift
  echo "== $label: $*"
  timeout ${RUN_TIMEOUT:-900} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2961$((RANDOM % 10)) \
    bench.py --gpus $N --warmup 5 "$@" > gpurun_out/${T}_$label.json 2> gpurun_out/${T}_$label.err
  python - <<PY
import json
for ln in open("gpurun_out/${T}_$label.json"):
    if ln.startswith("{"):
        d=json.loads(ln)
        print("$label value %.4g ms/step %.3f n_gpus %d" % (d["value"], d["ms_per_step"], d["n_gpus"]))
        c=d["config"]; print({k:c.get(k) for k in ("phase_ms_rank0","max_owned","max_local_with_ghosts","atoms_per_gpu","heat_current","final_T_K")})
        r=d.get("reference_gpu"); print("reference_gpu", r and (r.get("value"), r.get("path"), r.get("wall_s"), r.get("returncode")))
PY
  tail -3 gpurun_out/${T}_$label.err | cut -c1-300
}
if [ "$N" = "8" ]; then
  run c3_weak_n8 --steps 100
  run c4_unep_n8 --workload unep --steps 100
  run c3_strong_n8 --steps 200 --cells 25 --grid 2x2x2 --no-reference-gpu
  run c3_weak_blocks_n8 --steps 100 --grid 2x2x2 --no-reference-gpu
else
  run c5_si_n$N --workload si --steps 200
  [ -n "${WITH_C3:-}" ] && run c3_weak_n$N --steps 100
fi
ls gpurun_out | tail -20
