O=gpurun_out/r2f; mkdir -p $O
python __graft_entry__.py --smoke > $O/smoke.txt 2>&1; echo "smoke rc=$?" >> $O/smoke.txt
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.txt
timeout 600 python bench.py --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_reference.json 2>> $O/bench.err
timeout 300 python scripts/bench_needles.py > $O/needles.jsonl 2> $O/needles.err; echo "rc=$?" >> $O/needles.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu_launch_run.txt 2>&1
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_volume_ops.py -m gpu -q -k "needles or mounted" > $O/sanitizer_needles.txt 2>&1; echo "rc=$?" >> $O/sanitizer_needles.txt
tail -2 $O/smoke.txt; grep -E "passed|failed|rc=" $O/pytest_gpu.txt | tail -2; tail -1 $O/bench.err; cat $O/bench.json | cut -c1-400; cat $O/needles.jsonl; tail -2 $O/needles.err; grep -E "ERROR SUMMARY|passed|failed|rc=" $O/sanitizer_needles.txt | tail -3; grep -c rs10x4 $O/launches.csv
