# A/B timing of library variants inside ONE gpurun call (boxes differ by a few percent): scripts/libsnn_b200_<v>.bin
L=bindsnet_b200/csrc/libsnn_b200.so
cp $L /tmp/cur.so
run() { python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1 kernel_ms', round(d['roofline']['kernel_ms'],4), 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'rates', round(d.get('e2e_from_rates',{}).get('value',0)), 'launches', d['gpu_launches'])"; }
for v in "$@"; do cp scripts/libsnn_b200_$v.bin $L; run $v; SNN_B200_PROF=1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | tail -45 | grep -E "per step|slowest group|late-path cycles|early STDP|gather ahead|exchange read|gather\+neurons" | cut -c1-200; done
cp /tmp/cur.so $L
