cd /root/repo
timeout 900 python -m pytest tests -m gpu -q -rf --no-header 2>&1 | tail -15 > gpurun_out/r2b_tests5.txt
SNN_B200_GPROF=1 timeout 300 python scripts/generic_case.py metric 250 2 2> gpurun_out/gprof5_metric.txt
SNN_B200_GPROF=1 timeout 300 python scripts/generic_case.py c3 250 2 2> gpurun_out/gprof5_c3.txt
SNN_B200_GPROF=1 timeout 300 python scripts/c4_case.py 200 2 > gpurun_out/c4_case.txt 2> gpurun_out/gprof5_c4.txt
timeout 300 python bench.py --tier 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2e_generic_metric.json 2> gpurun_out/bench_r2e_generic_metric.err
timeout 300 python bench.py --config c3 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2e_c3.json 2> gpurun_out/bench_r2e_c3.err
timeout 300 python bench.py --config c4 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_r2e_c4.json 2> gpurun_out/bench_r2e_c4.err
timeout 400 ncu --set full --import-source on --clock-control none -k regex:snn_generic_window -s 1 -c 1 -o gpurun_out/ncu_r2e_generic_c4 -f python scripts/c4_case.py 40 2 > gpurun_out/ncu_r2e_generic_c4.log 2>&1
timeout 400 ncu --set full --import-source on --clock-control none -k regex:snn_generic_window -s 1 -c 1 -o gpurun_out/ncu_r2e_generic_metric -f python scripts/generic_case.py metric 100 2 > gpurun_out/ncu_r2e_generic_metric.log 2>&1
tail -4 gpurun_out/r2b_tests5.txt; cat gpurun_out/c4_case.txt
for f in metric c3 c4; do echo "== $f"; tail -9 gpurun_out/gprof5_$f.txt | cut -c17-; done
for f in generic_metric c3 c4; do python -c "import json,sys; d=json.load(open('gpurun_out/bench_r2e_$f.json')); print('$f', d['value'], d['ms_per_step'])"; done
