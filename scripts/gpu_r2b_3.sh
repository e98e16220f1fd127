cd /root/repo
timeout 900 python -m pytest tests -m gpu -q -rf --no-header 2>&1 | tail -15 > gpurun_out/r2b_tests3.txt
SNN_B200_GPROF=1 timeout 300 python scripts/generic_case.py metric 250 2 2> gpurun_out/gprof3_metric.txt
SNN_B200_GPROF=1 timeout 300 python scripts/generic_case.py c3 250 2 2> gpurun_out/gprof3_c3.txt
SNN_B200_GPROF=1 timeout 300 python bench.py --config c4 --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2> gpurun_out/gprof3_c4.txt
timeout 300 python bench.py --tier 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2c_generic_metric.json 2> gpurun_out/bench_r2c_generic_metric.err
timeout 300 python bench.py --config c3 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2c_c3.json 2> gpurun_out/bench_r2c_c3.err
timeout 300 python bench.py --config c4 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_r2c_c4.json 2> gpurun_out/bench_r2c_c4.err
tail -4 gpurun_out/r2b_tests3.txt; tail -9 gpurun_out/gprof3_metric.txt; tail -9 gpurun_out/gprof3_c3.txt; tail -9 gpurun_out/gprof3_c4.txt
for f in gpurun_out/bench_r2c_generic_metric.json gpurun_out/bench_r2c_c3.json gpurun_out/bench_r2c_c4.json; do python -c "import json,sys; d=json.load(open('$f')); print(d['config'].get('baseline_config'), d['value'], d['ms_per_step'])"; done
