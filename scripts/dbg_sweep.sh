# timing-only sweep of the fused kernel's switches (results are INVALID for SNN_B200_DEBUG != 0)
# usage: SWEEP="0 2 4" bash scripts/dbg_sweep.sh
for d in ${SWEEP:-0}; do
  SNN_B200_DEBUG=$d python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('dbg=$d kernel_ms', round(d['roofline']['kernel_ms'],3), 'us/step', round(d['roofline']['kernel_ms']*1000/250,2), 'value', round(d['value']), 'e2e', round(d['e2e']['value']))"
done
