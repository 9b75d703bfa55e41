set -x
for tool in memcheck racecheck synccheck; do
  timeout 500 compute-sanitizer --tool $tool --print-limit 10 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_san_smoke_$tool.log 2>&1
  tail -3 gpurun_out/r02_san_smoke_$tool.log
done
timeout 900 compute-sanitizer --tool memcheck --print-limit 10 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "compact or any_sketch or tie_path or gob or chunked" > gpurun_out/r02_san_tests_memcheck.log 2>&1
tail -4 gpurun_out/r02_san_tests_memcheck.log
