set -x
mkdir -p gpurun_out/r2k
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -12 > gpurun_out/r2k/pytest.log
timeout 600 python bench.py --config filip --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2k/bench_filip.json 2> gpurun_out/r2k/bench_filip.err
timeout 300 python tools/step_profile.py --config filip > gpurun_out/r2k/step_profile_filip.md 2>/dev/null
timeout 300 python tools/step_profile.py --config declip > gpurun_out/r2k/step_profile_declip.md 2>/dev/null
python tools/parity_report.py filip_vitb32_l12_b64 filip_vitb32_l2_b8 > gpurun_out/r2k/parity_filip.json 2>/dev/null
tail -4 gpurun_out/r2k/pytest.log
