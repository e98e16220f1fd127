cd /root/repo
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_abi.py -x -q 2>&1 | tail -5 > gpurun_out/r2_tests_c.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 >> gpurun_out/r2_tests_c.txt
