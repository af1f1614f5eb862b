set -x
mkdir -p gpurun_out/r2a
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2a/pytest.log
python tools/parity_report.py > gpurun_out/r2a/parity_report.json 2> gpurun_out/r2a/parity_report.err
for c in clip declip filip res50; do
  extra="--no-cpu-baseline"; [ $c = clip ] && extra=""
  timeout 600 python bench.py --config $c --steps 10 --warmup 3 $extra > gpurun_out/r2a/bench_$c.json 2> gpurun_out/r2a/bench_$c.err
  timeout 300 python tools/step_profile.py --config $c > gpurun_out/r2a/step_profile_$c.md 2> gpurun_out/r2a/step_profile_$c.err
done
timeout 600 python bench.py --config clip --head fused --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2a/bench_clip_fused.json 2> gpurun_out/r2a/bench_clip_fused.err
tail -5 gpurun_out/r2a/pytest.log
