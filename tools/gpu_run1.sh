set -x
nvidia-smi --query-gpu=name,memory.total --format=csv
nproc; free -g | head -2; lscpu | grep -E "Model name|NUMA|Socket" 
nvidia-smi topo -m 2>/dev/null | head -20
python -c "
from groot_b200 import api
print('int_issue_peak', api.int_issue_peak(0))
" > gpurun_out/r02_intpeak.txt 2>&1
cat gpurun_out/r02_intpeak.txt
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02_pytest_gpu_1.log
cat gpurun_out/r02_pytest_gpu_1.log
python tools/kernel_times.py 10000000 100 > gpurun_out/r02_kt_c3.txt 2>&1; tail -3 gpurun_out/r02_kt_c3.txt
python tools/kernel_times.py 10000000 150 > gpurun_out/r02_kt_c4.txt 2>&1; tail -3 gpurun_out/r02_kt_c4.txt
