set -x
mkdir -p gpurun_out/r2u
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_clip.py tests/test_gpu_resnet.py -q -x 2>&1 | tail -25 > gpurun_out/r2u/pytest.log
timeout 200 python tools/ln_perf.py > gpurun_out/r2u/ln_perf.txt 2>&1
DC_LN_BWD_GROUP=1 timeout 200 python tools/ln_perf.py >> gpurun_out/r2u/ln_perf.txt 2>&1
DC_LN_BWD_V1=1 timeout 200 python tools/ln_perf.py >> gpurun_out/r2u/ln_perf.txt 2>&1
for i in 1 2 3; do
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2> gpurun_out/r2u/e.txt | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('pipe', round(d['ms_per_step'],2), d['clocks']['sm_mhz'])" >> gpurun_out/r2u/ab.txt
  DC_LN_BWD_GROUP=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2> gpurun_out/r2u/e.txt | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('group', round(d['ms_per_step'],2), d['clocks']['sm_mhz'])" >> gpurun_out/r2u/ab.txt
done
timeout 600 python bench.py --config res50 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2u/bench_res50.json 2> gpurun_out/r2u/bench_res50.err
timeout 300 python tools/step_profile.py --config res50 > gpurun_out/r2u/step_profile_res50.md 2>/dev/null
tail -6 gpurun_out/r2u/pytest.log; cat gpurun_out/r2u/ln_perf.txt gpurun_out/r2u/ab.txt; grep resident gpurun_out/r2u/bench_res50.err
