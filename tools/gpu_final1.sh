set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r02_pytest_gpu.log
cat gpurun_out/r02_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
tail -c 300 gpurun_out/r02_bench.json; tail -2 gpurun_out/r02_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err
tail -c 300 gpurun_out/r02_bench_ref.json
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches.csv python tools/prof_step.py 10000000 100 compact > gpurun_out/r02_ncu_launches.log 2>&1
tail -1 gpurun_out/r02_ncu_launches.log
timeout 900 ncu --profile-from-start off --set full --clock-control none -k regex:'seed_kernel|fill_kernel|pack_reads|align_|project_' -f -o gpurun_out/r02_prof python tools/prof_step.py 2000000 100 compact > gpurun_out/r02_ncu_full.log 2>&1
tail -1 gpurun_out/r02_ncu_full.log
