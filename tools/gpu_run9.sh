set -x
timeout 300 python tools/e2e_trace.py 10000000 > gpurun_out/r02_e2e_trace.txt 2>&1
tail -30 gpurun_out/r02_e2e_trace.txt
