set -x
mkdir -p gpurun_out/r2g
N=${NGPU:-2}
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N --steps 12 --warmup 3 --no-e2e --no-cpu-baseline; }
v() { name=$1; shift; env "$@" bash -c "$(declare -f run); N=$N; run $PORT" > gpurun_out/r2g/$name.json 2> gpurun_out/r2g/$name.err; PORT=$((PORT+1)); }
PORT=29631
v a_1stream_buckets DECLIP_B200_TOWER_STREAMS=0 DECLIP_B200_NCCL_CTAS=0
v b_2stream_endbucket_bf16 DECLIP_B200_BUCKET_LAYERS=0 DECLIP_B200_NCCL_CTAS=0
v c_2stream_reserve4 DECLIP_B200_NCCL_CTAS=4
v d_2stream_maxctas8_noreserve DECLIP_B200_NCCL_CTAS=0 NCCL_MAX_CTAS=8
v e_2stream_buckets6 DECLIP_B200_BUCKET_LAYERS=6 DECLIP_B200_NCCL_CTAS=0
v f_1stream_r1style DECLIP_B200_TOWER_STREAMS=0 DECLIP_B200_BUCKET_LAYERS=0 DECLIP_B200_GRAD_DTYPE=fp32 DECLIP_B200_NCCL_CTAS=0
for f in gpurun_out/r2g/*.err; do echo $f; grep resident $f; done
