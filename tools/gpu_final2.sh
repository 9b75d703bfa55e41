set -x
timeout 500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py tests/test_gpu_cli.py -m gpu -x -q -k "compact or fixed_read or ranks or cli_oxa or travis or walk_variants or card" 2>&1 | tail -3
KT_COMPACT=1 timeout 200 python tools/kernel_times.py 10000000 100 2>&1 | tail -2
