run() { python bench.py --no-cpu-baseline --skip-e2e --parity-streams 0 --steps 40 2>/dev/null | python -c "
import json,sys
l=[json.loads(x) for x in sys.stdin if x.startswith('{')][0]
print('$1', round(l['ms_per_step'],1))"; }
for i in 1 2 3; do run default; done
export W2T_BENCH_NO_SAMPLER=1; for i in 1 2 3; do run nosampler; done; unset W2T_BENCH_NO_SAMPLER
export W2T_DBG_NO_CROWD16=1; for i in 1 2 3; do run no16; done
