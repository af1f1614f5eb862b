set -x
mkdir -p gpurun_out/r2l
N=${NGPU:-2}; export N
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
export -f run
DECLIP_B200_SYMM_HEAD=1 timeout 600 bash -c "run 29641 tools/dist_check.py --batch 256 --layers 1 --head fused" > gpurun_out/r2l/dist_check_symm_n$N.log 2>&1
tail -6 gpurun_out/r2l/dist_check_symm_n$N.log
timeout 600 bash -c "run 29642 tools/dist_check.py --batch 256 --layers 1 --head fused" > gpurun_out/r2l/dist_check_nccl_n$N.log 2>&1
tail -4 gpurun_out/r2l/dist_check_nccl_n$N.log
timeout 600 bash -c "run 29643 bench.py --gpus $N --steps 12 --warmup 3" > gpurun_out/r2l/bench_clip_n$N.json 2> gpurun_out/r2l/bench_clip_n$N.err
DECLIP_B200_SYMM_HEAD=1 timeout 600 bash -c "run 29644 bench.py --gpus $N --steps 12 --warmup 3 --no-e2e" > gpurun_out/r2l/bench_clip_symm_n$N.json 2> gpurun_out/r2l/bench_clip_symm_n$N.err
grep -h resident gpurun_out/r2l/*.err
