cd /root/repo
timeout 900 python -m pytest tests -m gpu -q -rf --no-header 2>&1 | tail -15 > gpurun_out/r2b_tests6.txt
SNN_B200_GPROF=1 timeout 300 python bench.py --config c4 --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2> gpurun_out/gprof6_c4.txt
timeout 300 python bench.py --config c4 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_r2f_c4.json 2> gpurun_out/bench_r2f_c4.err
timeout 300 python bench.py --tier 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2f_generic_metric.json 2> gpurun_out/bench_r2f_generic_metric.err
timeout 300 python bench.py --config c3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2f_c3.json 2> gpurun_out/bench_r2f_c3.err
timeout 400 ncu --set full --import-source on --clock-control none -k regex:snn_generic_window -s 1 -c 1 -o gpurun_out/ncu_r2f_generic_c4 -f python scripts/c4_case.py 40 2 > gpurun_out/ncu_r2f_generic_c4.log 2>&1
tail -4 gpurun_out/r2b_tests6.txt
tail -9 gpurun_out/gprof6_c4.txt | cut -c17-
for f in generic_metric c3 c4; do python -c "import json,sys; d=json.load(open('gpurun_out/bench_r2f_$f.json')); print('$f', d['value'], d['ms_per_step'])"; done
