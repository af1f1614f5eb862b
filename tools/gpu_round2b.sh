set -x
mkdir -p gpurun_out/r2b
timeout 900 python -m pytest tests/test_gpu_head.py "tests/test_gpu_ops.py::test_fused_adamw_state_dict_and_skipped_params" "tests/test_gpu_clip.py::test_fused_adamw_training_matches_torch_adamw" -q 2>&1 | grep -E "^E  |^tests/|^FAILED|passed|failed|Error" | cut -c1-400 > gpurun_out/r2b/pytest_fail.log
timeout 600 python tools/resnet_debug.py 8 > gpurun_out/r2b/resnet_debug.log 2>&1
BENCH_PER_STEP=1 timeout 600 python bench.py --config clip --head fused --steps 12 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2b/bench_clip_fused.json 2> gpurun_out/r2b/bench_clip_fused.err
timeout 300 python tools/step_profile.py --config clip --head fused > gpurun_out/r2b/step_profile_clip_fused.md 2> gpurun_out/r2b/step_profile_clip_fused.err
tail -20 gpurun_out/r2b/pytest_fail.log
