set -x
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gob" 2>&1 | tail -5
timeout 600 python bench.py --steps 10 --warmup 5 > gpurun_out/r02_bench_n1_a.json 2> gpurun_out/r02_bench_n1_a.err
tail -c 1500 gpurun_out/r02_bench_n1_a.json; tail -3 gpurun_out/r02_bench_n1_a.err
