set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "not c3_full and not full_size" 2>&1 | tail -3
timeout 200 python tools/kernel_times.py 10000000 100 2>&1 | tail -2
timeout 200 python tools/kernel_times.py 10000000 150 2>&1 | tail -2
