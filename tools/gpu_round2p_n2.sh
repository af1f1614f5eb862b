set -x
mkdir -p gpurun_out/r2p
N=2; export N
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
export -f run
DECLIP_B200_SYMM_HEAD=push timeout 600 bash -c "run 29661 tools/dist_check.py --batch 64 --layers 1 --head fused" > gpurun_out/r2p/dist_check_push_n2.log 2>&1
tail -5 gpurun_out/r2p/dist_check_push_n2.log
DECLIP_B200_SYMM_HEAD=push timeout 600 bash -c "run 29662 bench.py --gpus 2 --steps 12 --warmup 3 --no-e2e" > gpurun_out/r2p/bench_clip_push_n2.json 2> gpurun_out/r2p/bench_clip_push_n2.err
timeout 600 bash -c "run 29663 bench.py --gpus 2 --steps 12 --warmup 3 --no-e2e" > gpurun_out/r2p/bench_clip_nccl_n2.json 2> gpurun_out/r2p/bench_clip_nccl_n2.err
timeout 600 bash -c "run 29664 bench.py --gpus 2 --config declip --steps 8 --warmup 3 --no-e2e" > gpurun_out/r2p/bench_declip_n2.json 2> gpurun_out/r2p/bench_declip_n2.err
grep -h resident gpurun_out/r2p/*.err
