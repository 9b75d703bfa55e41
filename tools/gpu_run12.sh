set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
timeout 200 python tools/kernel_times.py 10000000 100 2>&1 | tail -2
timeout 300 python tools/e2e_trace.py 10000000 > gpurun_out/r02_e2e_trace.txt 2>&1
grep -E "^chunk|chunk [0-9]+ \(lane" gpurun_out/r02_e2e_trace.txt | tail -14
