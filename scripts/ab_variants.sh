for cfg in "0 0" "0 1" "0 2" "1 0"; do set -- $cfg; echo "TEAM=$1 VARIANT=$2"; B200MD_NEP_TEAM=$1 B200MD_NEP_VARIANT=$2 timeout 300 python bench.py --steps 60 --warmup 5 --no-cpu-baseline | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.4g'%d['value'], round(d['ms_per_step'],3), d['roofline']['stage_ms'])"; done
