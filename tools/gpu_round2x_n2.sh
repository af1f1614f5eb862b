set -x
mkdir -p gpurun_out/r2x
N=2; export N
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
export -f run
timeout 600 bash -c "run 29660 tools/dist_check.py --batch 64 --layers 1 --head fused" > gpurun_out/r2x/dist_check_n2.log 2>&1
tail -3 gpurun_out/r2x/dist_check_n2.log
port=29670
for rep in 1 2; do
for v in default overlap0 buckets0 nosync; do
  port=$((port+1))
  case $v in
    default) envs="";;
    overlap0) envs="DECLIP_B200_OVERLAP=0";;
    buckets0) envs="DECLIP_B200_BUCKET_LAYERS=0";;
    nosync) envs="DECLIP_B200_SKIP_GRAD_SYNC=1 DECLIP_B200_OVERLAP=0";;
  esac
  env $envs timeout 300 bash -c "run $port bench.py --gpus 2 --steps 16 --warmup 4 --no-e2e --no-cpu-baseline" > gpurun_out/r2x/b_$v.json 2> gpurun_out/r2x/b_$v.err
  python -c "import json; d=json.load(open('gpurun_out/r2x/b_$v.json')); print('$v', round(d['ms_per_step'],2), d['clocks']['sm_mhz'])" >> gpurun_out/r2x/ab.txt
done
done
timeout 300 python bench.py --steps 16 --warmup 4 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('n1', round(d['ms_per_step'],2), d['clocks']['sm_mhz'])" >> gpurun_out/r2x/ab.txt
cat gpurun_out/r2x/ab.txt
