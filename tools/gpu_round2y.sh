set -x
mkdir -p gpurun_out/r2y
timeout 300 python tools/step_jitter.py > gpurun_out/r2y/jitter.txt 2> gpurun_out/r2y/jitter.err
timeout 300 python -m pytest tests/test_gpu_filip.py tests/test_gpu_resnet.py -q 2>&1 | tail -5 > gpurun_out/r2y/pytest.log
timeout 600 python bench.py --config res50 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2y/bench_res50.json 2> gpurun_out/r2y/bench_res50.err
timeout 300 python tools/conv_shapes.py > gpurun_out/r2y/conv_shapes.md 2> gpurun_out/r2y/conv_shapes.err
cat gpurun_out/r2y/jitter.txt; tail -3 gpurun_out/r2y/pytest.log; tail -3 gpurun_out/r2y/jitter.err; grep resident gpurun_out/r2y/bench_res50.err; grep "igemm B" gpurun_out/r2y/conv_shapes.md | head -12
