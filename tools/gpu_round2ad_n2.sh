# 2 x B200 sanity of the final tree: dist parity check (fused head, bucketed gradients) + the N = 2 bench line
set -x
O=gpurun_out/r2ad; mkdir -p $O
N=2; export N
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
export -f run
timeout 240 bash -c "run 29660 tools/dist_check.py --batch 64 --layers 1 --head fused" > $O/dist_check_n2.log 2>&1
tail -3 $O/dist_check_n2.log
timeout 240 bash -c "run 29671 bench.py --gpus 2 --steps 16 --warmup 4 --no-cpu-baseline" > $O/bench_clip_n2.json 2> $O/bench_clip_n2.err
grep -h resident $O/*.err; cut -c1-300 $O/bench_clip_n2.json
