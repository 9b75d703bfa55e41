set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r02_pytest_gpu.log
cat gpurun_out/r02_pytest_gpu.log
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches.csv python tools/prof_step.py 10000000 100 compact > gpurun_out/r02_ncu_launches.log 2>&1
tail -1 gpurun_out/r02_ncu_launches.log
timeout 900 ncu --profile-from-start off --set full --clock-control none -k regex:'seed_kernel|fill_kernel|pack_reads|align_|project_' -f -o gpurun_out/r02_prof python tools/prof_step.py 2000000 100 compact > gpurun_out/r02_ncu_full.log 2>&1
tail -1 gpurun_out/r02_ncu_full.log
