O=gpurun_out/r2j; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "power_mode or roofline or encode_device" > $O/pytest.txt 2>&1; echo "rc=$?" >> $O/pytest.txt
for mode in 0 1 2; do SWEC_POWER_MODE=$mode timeout 300 python bench.py --batch-volumes 256 2>> $O/batch.err | sed "s/^/power_mode=$mode /" >> $O/batch256_modes.txt; done
timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > $O/bench_quick.json 2>> $O/batch.err
grep -E "passed|failed|rc=" $O/pytest.txt | tail -2; python - <<'PY'
import json
for l in open('gpurun_out/r2j/batch256_modes.txt'):
    m,_,j=l.partition(' '); d=json.loads(j); print(m, d['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks']['reasons'], d['clocks']['power_w_max'], d['digest'])
d=json.load(open('gpurun_out/r2j/bench_quick.json')); print('bench', d['value'], d['roofline']['frac'], d['reconstruct']['value'], d['reconstruct']['roofline_frac'])
PY
tail -2 $O/batch.err
