set -x
mkdir -p gpurun_out/r2v
timeout 300 python tools/conv_shapes.py > gpurun_out/r2v/conv_shapes.md 2> gpurun_out/r2v/conv_shapes.err
timeout 100 python tools/ln_perf.py > gpurun_out/r2v/ln_perf.txt 2>&1
DC_LN_BWD_GROUP=1 timeout 100 python tools/ln_perf.py >> gpurun_out/r2v/ln_perf.txt 2>&1
DC_LN_BWD_V1=1 timeout 100 python tools/ln_perf.py >> gpurun_out/r2v/ln_perf.txt 2>&1
tail -3 gpurun_out/r2v/conv_shapes.err; head -50 gpurun_out/r2v/conv_shapes.md; cat gpurun_out/r2v/ln_perf.txt | cut -c1-100
