set -x
nvidia-smi topo -m | head -8
python -m pytest tests/test_gpu_multi.py tests/test_gpu_cli.py -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r02_pytest_gpu_multi2.log
cat gpurun_out/r02_pytest_gpu_multi2.log
