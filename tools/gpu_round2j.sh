set -x
mkdir -p gpurun_out/r2j
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -12 > gpurun_out/r2j/pytest.log
for i in 1 2 3 4 5 6 7 8; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('run', round(d['ms_per_step'],2), d['clocks']['sm_mhz'], d['memory'])" >> gpurun_out/r2j/repeat.txt
done
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2j/bench_clip.json 2> gpurun_out/r2j/bench_clip.err
for c in declip filip res50; do timeout 600 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2j/bench_$c.json 2> gpurun_out/r2j/bench_$c.err; done
cat gpurun_out/r2j/repeat.txt; tail -3 gpurun_out/r2j/pytest.log
