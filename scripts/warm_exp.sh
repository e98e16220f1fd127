# diagnostic: does keeping the late path's code warm (every column group walks it every step, bit 256) change its cost?
run() { python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1 kernel_ms', round(d['roofline']['kernel_ms'],4), 'value', round(d['value']))"; }
prof() { SNN_B200_PROF=1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | tail -48 | grep -E "per step|slowest group|late-path cycles|late path, mean|winners|late set-up|late pass|gather\+neurons" | cut -c1-330; }
run normal; prof
export SNN_B200_DEBUG=256
run warm; prof
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "t40 or c2_" 2>&1 | tail -3
