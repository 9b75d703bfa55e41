set -x
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_cli.py -m gpu -x -q -k "fixed_read or cli_oxa or travis or chunked" 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
tail -c 200 gpurun_out/r02_bench.json; tail -2 gpurun_out/r02_bench.err
