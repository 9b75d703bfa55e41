set -x
timeout 600 python bench.py --steps 10 --warmup 5 > gpurun_out/r02_bench_n1_b.json 2> gpurun_out/r02_bench_n1_b.err
tail -c 400 gpurun_out/r02_bench_n1_b.json; tail -3 gpurun_out/r02_bench_n1_b.err
