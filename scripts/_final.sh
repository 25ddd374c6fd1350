O=gpurun_out/r2q; mkdir -p $O
python __graft_entry__.py --smoke > $O/smoke.txt 2>&1; echo "smoke rc=$?" >> $O/smoke.txt
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.txt
timeout 900 python bench.py --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_reference.json 2>> $O/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-files --no-sustained > $O/ncu_launch_run.txt 2>&1
tail -2 $O/smoke.txt; grep -E "passed|failed|rc=" $O/pytest_gpu.txt | tail -2; tail -1 $O/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2q/bench.json'))
print('value',d['value'],'frac',d['roofline']['frac'],'clocks',d['clocks']['sm_mhz'],d['clocks']['reasons'],'launches',d['gpu_launches'])
print('recon',d['reconstruct']['value'],d['reconstruct']['roofline_frac'],'sustained',d['sustained']['roofline_frac'],d['sustained']['reasons'])
print('e2e',d['e2e']['value'],d['e2e']['roofline']['frac'],'cpu',d['cpu_baseline']['value'])
print('files',d['file_level'])
print('ref',json.load(open('gpurun_out/r2q/bench_reference.json'))['value'])
PY
grep -c "rs10x4_encode" $O/launches.csv
