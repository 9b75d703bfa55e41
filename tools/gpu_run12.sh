set -x
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "c1 or device_projection or chunked or golden or no_align or stages" 2>&1 | tail -3
timeout 200 python tools/kernel_times.py 10000000 100 2>&1 | tail -2
