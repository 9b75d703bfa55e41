set -x
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r02_pytest_gpu_multi2.log
cat gpurun_out/r02_pytest_gpu_multi2.log
timeout 300 python -m pytest tests/test_gpu_cli.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02_pytest_gpu_cli2.log
cat gpurun_out/r02_pytest_gpu_cli2.log
timeout 240 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "any_sketch or compact" 2>&1 | tail -15 > gpurun_out/r02_pytest_gpu_generic.log
cat gpurun_out/r02_pytest_gpu_generic.log
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 2 --reads 4000000 --extra none > gpurun_out/r02_bench_n2_smoke.json 2> gpurun_out/r02_bench_n2_smoke.err
tail -c 3000 gpurun_out/r02_bench_n2_smoke.json; tail -5 gpurun_out/r02_bench_n2_smoke.err
