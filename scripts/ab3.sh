# A/B of library variants (scripts/libsnn_b200_<v>.bin) in one gpurun call + instruction-cache counters of each
L=bindsnet_b200/csrc/libsnn_b200.so
cp $L /tmp/cur.so
M=gcc__cache_requests_type_instruction,gcc__cache_requests_type_instruction_lookup_miss,gcc__gcc2xbar_requests_type_instruction,gcc__cache_requests_type_constant,gcc__cache_requests_type_constant_lookup_miss,gcc__gcc2xbar_requests_type_constant,smsp__warps_issue_stalled_no_instruction,smsp__warps_issue_stalled_imc_miss,smsp__warps_active,gpu__time_duration.sum
run() { python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1 kernel_ms', round(d['roofline']['kernel_ms'],4), 'value', round(d['value']))"; }
prof() { SNN_B200_PROF=1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | tail -48 | grep -E "per step|slowest group|late-path cycles|late path, mean" | cut -c1-330; }
for v in "$@"; do cp scripts/libsnn_b200_$v.bin $L; run $v; prof
  if [ -n "$NCU" ]; then timeout 300 ncu --metrics $M -k regex:snn_dc2_window --launch-skip 3 --launch-count 1 --clock-control none python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2>/dev/null | grep -E "gcc__|smsp__|gpu__time" | sed 's/  */ /g'; fi
done
cp /tmp/cur.so $L
