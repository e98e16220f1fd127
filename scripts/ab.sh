# A/B/C timing of library variants inside ONE gpurun call (boxes differ by a few percent)
L=bindsnet_b200/csrc/libsnn_b200.so
cp $L /tmp/A.so
run() { python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1 kernel_ms', round(d['roofline']['kernel_ms'],4), 'value', round(d['value']), 'e2e', round(d['e2e']['value']))"; }
run A
for v in "$@"; do cp scripts/libsnn_b200_$v.bin $L; run $v; done
cp /tmp/A.so $L; run A
