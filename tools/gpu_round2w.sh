set -x
mkdir -p gpurun_out/r2w
timeout 900 python -m pytest tests/test_gpu_resnet.py tests/test_gpu_fullsize.py tests/test_gpu_slip.py -q -x 2>&1 | tail -25 > gpurun_out/r2w/pytest.log
timeout 600 python bench.py --config res50 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2w/bench_res50.json 2> gpurun_out/r2w/bench_res50.err
timeout 300 python tools/conv_shapes.py > gpurun_out/r2w/conv_shapes.md 2> gpurun_out/r2w/conv_shapes.err
timeout 300 python tools/step_profile.py --config res50 > gpurun_out/r2w/step_profile_res50.md 2>/dev/null
tail -6 gpurun_out/r2w/pytest.log; grep resident gpurun_out/r2w/bench_res50.err
