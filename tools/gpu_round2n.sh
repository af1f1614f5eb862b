set -x
mkdir -p gpurun_out/r2n
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2n/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2n/smoke.log 2>&1
timeout 600 python bench.py --config declip --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2n/bench_declip.json 2> gpurun_out/r2n/bench_declip.err
timeout 600 python bench.py --config declip --head strips --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2n/bench_declip_strips.json 2> gpurun_out/r2n/bench_declip_strips.err
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2n/bench_clip.json 2> gpurun_out/r2n/bench_clip.err
tail -4 gpurun_out/r2n/pytest.log; cat gpurun_out/r2n/smoke.log | tail -2
