set -x
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -4
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 5 > gpurun_out/r02_bench_n2_a.json 2> gpurun_out/r02_bench_n2_a.err
tail -c 600 gpurun_out/r02_bench_n2_a.json; tail -3 gpurun_out/r02_bench_n2_a.err
