set -x
mkdir -p gpurun_out/r2e
N=${NGPU:-2}
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
timeout 600 bash -c "$(declare -f run); N=$N; run 29621 tools/dist_check.py --batch 32 --layers 2" > gpurun_out/r2e/dist_check_n$N.log 2>&1
tail -8 gpurun_out/r2e/dist_check_n$N.log
timeout 600 bash -c "$(declare -f run); N=$N; run 29622 bench.py --gpus $N --steps 10 --warmup 3" > gpurun_out/r2e/bench_clip_n$N.json 2> gpurun_out/r2e/bench_clip_n$N.err
DECLIP_B200_BUCKET_LAYERS=0 DECLIP_B200_GRAD_DTYPE=fp32 DECLIP_B200_NCCL_CTAS=0 timeout 600 bash -c "$(declare -f run); N=$N; run 29623 bench.py --gpus $N --steps 10 --warmup 3 --no-e2e" > gpurun_out/r2e/bench_clip_n${N}_r1style.json 2> gpurun_out/r2e/bench_clip_n${N}_r1style.err
DECLIP_B200_NCCL_CTAS=0 timeout 600 bash -c "$(declare -f run); N=$N; run 29624 bench.py --gpus $N --steps 10 --warmup 3 --no-e2e" > gpurun_out/r2e/bench_clip_n${N}_noreserve.json 2> gpurun_out/r2e/bench_clip_n${N}_noreserve.err
tail -2 gpurun_out/r2e/*.err | tail -30
