cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2_tests_a.txt
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_v1sh.json 2> gpurun_out/bench_r2_v1sh.err
python bench.py --tier 3 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r2_v2sh.json 2>/dev/null
cp bindsnet_b200/csrc/libsnn_b200.so /tmp/cur.so; cp scripts/libsnn_b200_base.bin bindsnet_b200/csrc/libsnn_b200.so
python bench.py --tier 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r2_v1old.json 2>/dev/null
cp /tmp/cur.so bindsnet_b200/csrc/libsnn_b200.so
timeout 600 python bench.py --config c3 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2_c3_generic.json 2> gpurun_out/bench_r2_c3.err
