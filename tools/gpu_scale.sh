set -x
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/r02_bench_n2_final.json 2> gpurun_out/r02_bench_n2_final.err
tail -c 300 gpurun_out/r02_bench_n2_final.json; tail -3 gpurun_out/r02_bench_n2_final.err
