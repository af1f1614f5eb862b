set -x
mkdir -p gpurun_out/r2q
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ln_bwd -s 3 -c 1 -o gpurun_out/r2q/ln_bwd -f python tools/kernels_one.py > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_tc_kernel -s 6 -c 2 -o gpurun_out/r2q/attn -f python tools/kernels_one.py > /dev/null 2>&1
ls -la gpurun_out/r2q
