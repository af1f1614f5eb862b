set -x
mkdir -p gpurun_out/r2c
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > gpurun_out/r2c/pytest.log
python tools/parity_report.py clip_res50_l3463_b32 clip_res50_l1111_b4 declip_vitb32_l12_b64 > gpurun_out/r2c/parity_report.json 2> gpurun_out/r2c/parity_report.err
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c/bench_clip.json 2> gpurun_out/r2c/bench_clip.err
DECLIP_B200_TOWER_STREAMS=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2c/bench_clip_1stream.json 2> gpurun_out/r2c/bench_clip_1stream.err
timeout 300 python tools/gemm_perf.py > gpurun_out/r2c/gemm_perf.jsonl 2>&1
tail -5 gpurun_out/r2c/pytest.log
