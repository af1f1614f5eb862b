set -x
mkdir -p gpurun_out/r2m
N=8; export N
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
export -f run
timeout 900 bash -c "run 29651 tools/dist_check.py --batch 64 --layers 3" > gpurun_out/r2m/dist_check_n8.log 2>&1
tail -8 gpurun_out/r2m/dist_check_n8.log
timeout 600 bash -c "run 29652 bench.py --gpus 8 --steps 12 --warmup 3" > gpurun_out/r2m/bench_clip_n8.json 2> gpurun_out/r2m/bench_clip_n8.err
DECLIP_B200_BUCKET_LAYERS=0 timeout 600 bash -c "run 29653 bench.py --gpus 8 --steps 12 --warmup 3 --no-e2e" > gpurun_out/r2m/bench_clip_n8_endbucket.json 2> gpurun_out/r2m/bench_clip_n8_endbucket.err
DECLIP_B200_BUCKET_LAYERS=3 timeout 600 bash -c "run 29654 bench.py --gpus 8 --steps 12 --warmup 3 --no-e2e" > gpurun_out/r2m/bench_clip_n8_bucket3.json 2> gpurun_out/r2m/bench_clip_n8_bucket3.err
if [ -n "$TRY_SYMM" ]; then DECLIP_B200_SYMM_HEAD=1 timeout 600 bash -c "run 29655 bench.py --gpus 8 --steps 12 --warmup 3 --no-e2e" > gpurun_out/r2m/bench_clip_n8_symm.json 2> gpurun_out/r2m/bench_clip_n8_symm.err; fi
timeout 600 bash -c "run 29656 bench.py --gpus 8 --config declip --steps 8 --warmup 3 --no-e2e" > gpurun_out/r2m/bench_declip_n8.json 2> gpurun_out/r2m/bench_declip_n8.err
grep -h resident gpurun_out/r2m/*.err
