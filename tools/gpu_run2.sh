set -x
python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r02_pytest_gpu_2.log
cat gpurun_out/r02_pytest_gpu_2.log
python tools/kernel_times.py 10000000 100 > gpurun_out/r02_kt_c3_b.txt 2>&1; tail -3 gpurun_out/r02_kt_c3_b.txt
