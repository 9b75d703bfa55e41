set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw,temperature.gpu,clocks_event_reasons.active --format=csv
(for i in 1 2 3 4 5 6 7 8 9 10 11 12; do nvidia-smi --query-gpu=clocks.sm,power.draw,clocks_event_reasons.active --format=csv,noheader; sleep 2; done) > gpurun_out/clk.txt &
timeout 200 python tools/kernel_times.py 10000000 100 2>&1 | tail -2
wait
sort gpurun_out/clk.txt | uniq -c
