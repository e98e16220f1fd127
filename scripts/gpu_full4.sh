cd /root/repo
# round-2 evidence for the metric's kernel (tier 2): full ncu capture, launch list; c4 bench; reference arm
timeout 500 ncu --set full --import-source on --clock-control none -k regex:snn_dc_fused_window --launch-skip 3 --launch-count 1 -f -o gpurun_out/prof_r2_fused_v1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_r2_v1.log 2>&1
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2_fused_v1.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
python bench.py --config c4 --steps 3 --warmup 2 > gpurun_out/bench_r2_c4_generic.json 2> gpurun_out/bench_r2_c4.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r2_reference_arm.json 2> gpurun_out/bench_r2_reference_arm.err
python bench.py > gpurun_out/bench_r2_default.json 2> gpurun_out/bench_r2_default.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.txt 2>&1
