set -x
mkdir -p gpurun_out/r2f
for rep in 1 2; do
  DECLIP_B200_LIB=$PWD/declip_b200/_C_bufs2.so timeout 300 python tools/gemm_perf.py > gpurun_out/r2f/gemm_perf_bufs2_$rep.jsonl 2>&1
  timeout 300 python tools/gemm_perf.py > gpurun_out/r2f/gemm_perf_bufs4_$rep.jsonl 2>&1
done
DECLIP_B200_LIB=$PWD/declip_b200/_C_bufs2.so timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2f/bench_bufs2.json 2>/dev/null
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2f/bench_bufs4.json 2>/dev/null
DECLIP_B200_LIB=$PWD/declip_b200/_C_bufs2.so timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2f/bench_bufs2_b.json 2>/dev/null
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2f/bench_bufs4_b.json 2>/dev/null
timeout 300 python tools/gpu_eager_baseline.py 512 6 > gpurun_out/r2f/gpu_eager_baseline.json 2> gpurun_out/r2f/gpu_eager_baseline.err
timeout 300 python tools/e2e_strings.py 10 > gpurun_out/r2f/e2e_strings.json 2> gpurun_out/r2f/e2e_strings.err
python -m pytest tests/test_gpu_ops.py -q -x 2>&1 | tail -3
