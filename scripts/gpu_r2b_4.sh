cd /root/repo
timeout 900 python -m pytest tests -m gpu -q -rf --no-header 2>&1 | tail -15 > gpurun_out/r2b_tests4.txt
SNN_B200_GPROF=1 timeout 300 python scripts/generic_case.py metric 250 2 2> gpurun_out/gprof4_metric.txt
SNN_B200_GPROF=1 SNN_B200_GVAR=2 timeout 300 python scripts/generic_case.py metric 250 2 2> gpurun_out/gprof4_metric_v2.txt
SNN_B200_GPROF=1 timeout 300 python scripts/generic_case.py c3 250 2 2> gpurun_out/gprof4_c3.txt
SNN_B200_GPROF=1 timeout 300 python bench.py --config c4 --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2> gpurun_out/gprof4_c4.txt
SNN_B200_GPROF=1 SNN_B200_GVAR=2 timeout 300 python bench.py --config c4 --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2> gpurun_out/gprof4_c4_v2.txt
timeout 300 python bench.py --tier 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2d_generic_metric.json 2> gpurun_out/bench_r2d_generic_metric.err
SNN_B200_GVAR=2 timeout 300 python bench.py --tier 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2d_generic_metric_v2.json 2>/dev/null
timeout 300 python bench.py --config c3 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2d_c3.json 2> gpurun_out/bench_r2d_c3.err
timeout 300 python bench.py --config c4 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_r2d_c4.json 2> gpurun_out/bench_r2d_c4.err
SNN_B200_GVAR=2 timeout 300 python bench.py --config c4 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_r2d_c4_v2.json 2>/dev/null
tail -4 gpurun_out/r2b_tests4.txt
for f in metric metric_v2 c3 c4 c4_v2; do echo "== $f"; tail -9 gpurun_out/gprof4_$f.txt | cut -c17-; done
for f in generic_metric generic_metric_v2 c3 c4 c4_v2; do python -c "import json,sys; d=json.load(open('gpurun_out/bench_r2d_$f.json')); print('$f', d['value'], d['ms_per_step'])"; done
