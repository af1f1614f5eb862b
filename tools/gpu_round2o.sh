set -x
mkdir -p gpurun_out/r2o
timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_declip.py tests/test_gpu_defilip.py -q -x 2>&1 | tail -8 > gpurun_out/r2o/pytest.log
timeout 600 python bench.py --config declip --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2o/bench_declip.json 2> gpurun_out/r2o/bench_declip.err
timeout 300 python tools/step_profile.py --config declip > gpurun_out/r2o/step_profile_declip.md 2>/dev/null
tail -3 gpurun_out/r2o/pytest.log; grep resident gpurun_out/r2o/bench_declip.err
