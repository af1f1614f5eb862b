set -x
mkdir -p gpurun_out/r2d
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > gpurun_out/r2d/pytest.log
timeout 300 python tools/gemm_perf.py > gpurun_out/r2d/gemm_perf.jsonl 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2d/bench_clip.json 2> gpurun_out/r2d/bench_clip.err
timeout 300 python tools/step_profile.py --config clip > gpurun_out/r2d/step_profile_clip.md 2>/dev/null
tail -5 gpurun_out/r2d/pytest.log
