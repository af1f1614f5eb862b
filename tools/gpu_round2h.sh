set -x
mkdir -p gpurun_out/r2h
for i in 1 2 3 4 5; do
  DC_PDL=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('pdl0', round(d['ms_per_step'],2), d['clocks']['sm_mhz'])" >> gpurun_out/r2h/pdl_ab.txt
  DC_PDL=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('pdl1', round(d['ms_per_step'],2), d['clocks']['sm_mhz'])" >> gpurun_out/r2h/pdl_ab.txt
done
cat gpurun_out/r2h/pdl_ab.txt
