cd /root/repo
timeout 900 python -m pytest tests -m gpu -q -rf --no-header 2>&1 | tail -15 > gpurun_out/r2b_tests7.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke7.txt 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2g_default20.json 2> gpurun_out/bench_r2g_default20.err
timeout 300 python bench.py --config c4 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_r2g_c4.json 2> gpurun_out/bench_r2g_c4.err
timeout 300 python bench.py --config c3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2g_c3.json 2> gpurun_out/bench_r2g_c3.err
timeout 300 python bench.py --tier 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2g_generic_metric.json 2> gpurun_out/bench_r2g_generic_metric.err
tail -4 gpurun_out/r2b_tests7.txt; tail -2 gpurun_out/smoke7.txt
for f in default20 generic_metric c3 c4; do python -c "import json,sys; d=json.load(open('gpurun_out/bench_r2g_$f.json')); print('$f', d['value'], d['ms_per_step'], d.get('e2e',{}).get('value'), d['roofline'].get('frac'))"; done
