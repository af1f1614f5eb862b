# re-entry check of the committed tree: every GPU test, smoke(), the default bench, res50 after the 3-CTA conv change
set -x
O=gpurun_out/r2z; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x --durations=8 2>&1 | tail -25 > $O/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > $O/bench_clip.json 2> $O/bench_clip.err
timeout 300 python bench.py --config res50 --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_res50.json 2> $O/bench_res50.err
tail -14 $O/pytest.log; tail -2 $O/smoke.log; grep -h resident $O/*.err
