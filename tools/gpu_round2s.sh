set -x
mkdir -p gpurun_out/r2s
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_clip.py -q -x 2>&1 | tail -12 > gpurun_out/r2s/pytest.log
python tools/ln_perf.py > gpurun_out/r2s/ln_perf_v2.txt 2>&1
DC_LN_BWD_V1=1 python tools/ln_perf.py > gpurun_out/r2s/ln_perf_v1.txt 2>&1
for i in 1 2 3; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('v2', round(d['ms_per_step'],2), d['clocks']['sm_mhz'])" >> gpurun_out/r2s/ab.txt
  DC_LN_BWD_V1=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('v1', round(d['ms_per_step'],2), d['clocks']['sm_mhz'])" >> gpurun_out/r2s/ab.txt
done
tail -3 gpurun_out/r2s/pytest.log; cat gpurun_out/r2s/ln_perf_v2.txt gpurun_out/r2s/ln_perf_v1.txt gpurun_out/r2s/ab.txt
