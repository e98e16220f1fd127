cd /root/repo
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q --no-header -k "nccl" 2>&1 | tail -5 > gpurun_out/r2b_tests8_2gpu.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r2g_n2.json 2> gpurun_out/bench_r2g_n2.err
tail -3 gpurun_out/r2b_tests8_2gpu.txt
python -c "import json; d=json.load(open('gpurun_out/bench_r2g_n2.json')); print('n2', d['value'], d['ms_per_step'], d.get('e2e',{}).get('value'), d['n_gpus'])"
tail -3 gpurun_out/bench_r2g_n2.err
