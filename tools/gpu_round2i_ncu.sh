set -x
mkdir -p gpurun_out/r2i
python tools/head_one.py 512 8 512 5 > gpurun_out/r2i/head_one_w8.txt 2>&1
python tools/head_one.py 512 1 512 5 >> gpurun_out/r2i/head_one_w8.txt 2>&1
# (1) launch list of one bench step (serialised, cold cache: shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 2100 --launch-count 720 --csv --log-file gpurun_out/r2i/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2i/launches_bench.log 2>&1
# (2) full captures
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm2 -s 5 -c 1 -o gpurun_out/r2i/gemm_gelu -f python tools/gemm_one.py 1 25600 3072 768 gelu > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:head_fwd -s 2 -c 1 -o gpurun_out/r2i/head_fwd -f python tools/head_one.py 512 8 512 4 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:head_bwd -s 2 -c 1 -o gpurun_out/r2i/head_bwd -f python tools/head_one.py 512 8 512 4 > /dev/null 2>&1
ls -la gpurun_out/r2i
