# final validation of the tree after the attention-forward / Conv1x1Skip changes: every GPU test, smoke(), the four bench
# configs, the reference arm, the launch list and the in-situ step profile
set -x
O=gpurun_out/r2final2; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 > $O/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_clip.json 2> $O/bench_clip.err
timeout 300 python bench.py --config declip --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_declip.json 2> $O/bench_declip.err
timeout 300 python bench.py --config filip --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_filip.json 2> $O/bench_filip.err
timeout 300 python bench.py --config res50 --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_res50.json 2> $O/bench_res50.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
timeout 300 python tools/step_profile.py --config clip > $O/step_profile_clip.md 2>/dev/null
timeout 300 python tools/step_profile.py --config res50 > $O/step_profile_res50.md 2>/dev/null
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/launches_clip.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > $O/ncu_bench.log 2>&1
tail -5 $O/pytest.log; tail -2 $O/smoke.log; grep -h resident $O/*.err
