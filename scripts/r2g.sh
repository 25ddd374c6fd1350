#!/bin/bash
# round-2 session G (1 GPU): final validation of the last changes + two short probes
OUT=gpurun_out/r2g; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.txt; tail -3 $OUT/pytest_gpu.txt
timeout 200 python scripts/bench_needles.py > $OUT/needles.jsonl 2> $OUT/needles.err; cat $OUT/needles.jsonl
timeout 200 python scripts/experiments/batch_gap_rows.py > $OUT/batch_gap_rows.jsonl 2> $OUT/batch_gap_rows.err; cat $OUT/batch_gap_rows.jsonl; tail -2 $OUT/batch_gap_rows.err
timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" >> $OUT/bench.err; tail -2 $OUT/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2g/bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'frac',d['roofline']['frac'],'recon',d['reconstruct']['roofline_frac'],d['reconstruct']['single_loss']['roofline_frac'])
print('sustained',d['sustained']['roofline_frac'],d['sustained']['low_power_variant_from_step'],'batch',d['batch']['roofline_frac'],d['batch']['digest'],'e2e',d['e2e']['value'])
print('files',d['file_level'])
PY
