cd /root/repo
timeout 900 python -m pytest tests -m gpu -q -rf --no-header 2>&1 | tail -40 > gpurun_out/r2b_tests1.txt
timeout 300 python bench.py --tier 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2b_generic_metric.json 2> gpurun_out/bench_r2b_generic_metric.err
timeout 300 python bench.py --config c3 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2b_c3.json 2> gpurun_out/bench_r2b_c3.err
timeout 300 python bench.py --config c4 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_r2b_c4.json 2> gpurun_out/bench_r2b_c4.err
tail -3 gpurun_out/r2b_tests1.txt; head -c 400 gpurun_out/bench_r2b_generic_metric.json; echo; head -c 300 gpurun_out/bench_r2b_c3.json; echo; head -c 300 gpurun_out/bench_r2b_c4.json
