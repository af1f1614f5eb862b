# attention: two-CTAs-per-SM forward + split-barrier backward vs the round-1 kernels (DC_ATTN_FWD_V1 / DC_ATTN_BWD_V1)
set -x
O=gpurun_out/r2aa; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_ops.py -q -x -k attention 2>&1 | tail -5 > $O/pytest_attn.log
timeout 300 python tools/attn_probe.py --perf > $O/probe_new.txt 2>&1
DC_ATTN_FWD_V1=1 DC_ATTN_BWD_V1=1 timeout 300 python tools/attn_probe.py --perf > $O/probe_old.txt 2>&1
tail -3 $O/pytest_attn.log; grep -h "ALL_OK\|MISMATCH\|perf" $O/probe_new.txt $O/probe_old.txt
if grep -q ALL_OK $O/probe_new.txt; then
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > $O/pytest.log
timeout 300 python bench.py --steps 16 --warmup 4 --no-cpu-baseline > $O/bench_clip_new.json 2> $O/bench_clip_new.err
DC_ATTN_FWD_V1=1 DC_ATTN_BWD_V1=1 timeout 300 python bench.py --steps 16 --warmup 4 --no-cpu-baseline > $O/bench_clip_old.json 2> $O/bench_clip_old.err
timeout 300 python bench.py --steps 16 --warmup 4 --no-cpu-baseline > $O/bench_clip_new2.json 2> $O/bench_clip_new2.err
DC_ATTN_BWD_V1=1 timeout 300 python bench.py --steps 16 --warmup 4 --no-cpu-baseline > $O/bench_clip_fwdonly.json 2> $O/bench_clip_fwdonly.err
tail -4 $O/pytest.log; grep -H resident $O/*.err
fi
