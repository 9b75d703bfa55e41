set -x
nvidia-smi topo -m | head -12
nproc; free -g | head -2
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 5 > gpurun_out/r02_bench_n8_a.json 2> gpurun_out/r02_bench_n8_a.err
tail -c 600 gpurun_out/r02_bench_n8_a.json; tail -3 gpurun_out/r02_bench_n8_a.err
timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "8 or 4" 2>&1 | tail -4 > gpurun_out/r02_pytest_gpu_multi8.log
cat gpurun_out/r02_pytest_gpu_multi8.log
