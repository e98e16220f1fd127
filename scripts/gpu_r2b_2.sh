cd /root/repo
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --no-header -k "boosted or mcp or local or mstdpet" 2>&1 | tail -5 > gpurun_out/r2b_tests2.txt
SNN_B200_GPROF=1 timeout 300 python scripts/generic_case.py metric 250 2 2> gpurun_out/gprof_metric.txt
SNN_B200_GPROF=1 timeout 300 python scripts/generic_case.py c3 250 2 2> gpurun_out/gprof_c3.txt
SNN_B200_GPROF=1 timeout 300 python bench.py --config c4 --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2> gpurun_out/gprof_c4.txt
timeout 600 ncu --set full --import-source on --clock-control none -k regex:snn_generic_window -s 1 -c 1 -o gpurun_out/ncu_r2b_generic_metric -f python scripts/generic_case.py metric 60 2 > gpurun_out/ncu_r2b_generic_metric.log 2>&1
tail -3 gpurun_out/r2b_tests2.txt; tail -9 gpurun_out/gprof_metric.txt; tail -9 gpurun_out/gprof_c3.txt; tail -9 gpurun_out/gprof_c4.txt; tail -2 gpurun_out/ncu_r2b_generic_metric.log
