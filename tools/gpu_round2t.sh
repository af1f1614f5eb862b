set -x
mkdir -p gpurun_out/r2t
timeout 900 python -m pytest tests/test_gpu_resnet.py tests/test_gpu_fullsize.py tests/test_gpu_slip.py -q -x 2>&1 | tail -25 > gpurun_out/r2t/pytest.log
timeout 600 python bench.py --config res50 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2t/bench_res50.json 2> gpurun_out/r2t/bench_res50.err
DECLIP_B200_CONV_WGRAD=im2col timeout 600 python bench.py --config res50 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2t/bench_res50_im2colwgrad.json 2> gpurun_out/r2t/bench_res50_im2colwgrad.err
timeout 300 python tools/step_profile.py --config res50 > gpurun_out/r2t/step_profile_res50.md 2>/dev/null
tail -6 gpurun_out/r2t/pytest.log; grep resident gpurun_out/r2t/bench_res50.err gpurun_out/r2t/bench_res50_im2colwgrad.err
