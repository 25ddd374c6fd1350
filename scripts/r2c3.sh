#!/bin/bash
# round-2 session C3: full GPU test suite after the power-policy fix, default bench line, policy probe
OUT=gpurun_out/r2c3; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.txt; tail -4 $OUT/pytest_gpu.txt
timeout 120 python scripts/experiments/power_policy_probe.py > $OUT/power_policy.txt 2>&1; grep -E "==|swec\]" $OUT/power_policy.txt | head -12
timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" >> $OUT/bench.err; tail -2 $OUT/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c3/bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'frac',d['roofline']['frac'],'ms',d['ms_per_step'])
print('recon',d['reconstruct']['roofline_frac'],'sustained',d['sustained']['roofline_frac'],d['sustained']['ms_every_10th_step'])
print('variant',d['variant_30000MiB']['roofline_frac'],'batch',d['batch']['roofline_frac'],d['batch']['digest'])
print('e2e',d['e2e']['value'],'files',d['file_level'])
PY
