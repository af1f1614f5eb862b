# attention forward: two CTAs per SM (DC_ATTN_FWD2=1) vs the one-tile-per-SM kernel (default)
set -x
O=gpurun_out/r2ac; mkdir -p $O
DC_ATTN_FWD2=1 timeout 300 python -m pytest tests/test_gpu_ops.py -q -x -k attention 2>&1 | tail -5 > $O/pytest_attn.log
DC_ATTN_FWD2=1 timeout 300 python tools/attn_probe.py --perf > $O/probe_fwd2.txt 2>&1
timeout 300 python tools/attn_probe.py --perf > $O/probe_v1.txt 2>&1
tail -3 $O/pytest_attn.log; grep -h "ALL_OK\|MISMATCH\|perf" $O/probe_fwd2.txt $O/probe_v1.txt
if grep -q ALL_OK $O/probe_fwd2.txt; then
DC_ATTN_FWD2=1 timeout 300 python bench.py --steps 16 --warmup 4 --no-cpu-baseline --no-e2e > $O/bench_clip_fwd2.json 2> $O/bench_clip_fwd2.err
timeout 300 python bench.py --steps 16 --warmup 4 --no-cpu-baseline --no-e2e > $O/bench_clip_v1.json 2> $O/bench_clip_v1.err
DC_ATTN_FWD2=1 timeout 300 python bench.py --steps 16 --warmup 4 --no-cpu-baseline --no-e2e > $O/bench_clip_fwd2b.json 2> $O/bench_clip_fwd2b.err
timeout 300 python bench.py --steps 16 --warmup 4 --no-cpu-baseline --no-e2e > $O/bench_clip_v1b.json 2> $O/bench_clip_v1b.err
grep -H resident $O/*.err
DC_ATTN_FWD2=1 timeout 200 ncu --metrics gpu__time_duration.sum,launch__grid_size,launch__waves_per_multiprocessor --clock-control none -k regex:attn_tc_fwd2 -s 2 -c 2 --csv --log-file $O/ncu_fwd2.csv python tools/attn_one.py vit > /dev/null 2>&1
DC_ATTN_FWD2=1 timeout 200 ncu --metrics gpu__time_duration.sum,launch__grid_size,launch__waves_per_multiprocessor --clock-control none -k regex:attn_tc_fwd2 -s 2 -c 2 --csv --log-file $O/ncu_fwd2_text.csv python tools/attn_one.py text > /dev/null 2>&1
tail -3 $O/ncu_fwd2.csv $O/ncu_fwd2_text.csv
fi
# ResNet: skip-branch gradient added in conv1's dgrad epilogue (Conv1x1Skip)
timeout 600 python -m pytest tests/test_gpu_resnet.py tests/test_gpu_fullsize.py -q -x 2>&1 | tail -5 > $O/pytest_resnet.log
timeout 300 python bench.py --config res50 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $O/bench_res50.json 2> $O/bench_res50.err
tail -3 $O/pytest_resnet.log; grep -H resident $O/bench_res50.err
