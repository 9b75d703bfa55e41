set -x
timeout 700 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r02_pytest_gpu_7.log
cat gpurun_out/r02_pytest_gpu_7.log
for v in "" _mb5; do
  GROOTGPU_LIB=$PWD/groot_b200/libgrootgpu$v.so timeout 200 python tools/kernel_times.py 10000000 100 > gpurun_out/r02_kt7_c3$v.txt 2>&1; tail -2 gpurun_out/r02_kt7_c3$v.txt
done
