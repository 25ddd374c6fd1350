#!/bin/bash
# round-2 session E (1 GPU): final-state validation — GPU suite, smoke, default bench line + reference arm, file benches
OUT=gpurun_out/r2e; mkdir -p $OUT
python __graft_entry__.py --smoke > $OUT/smoke.txt 2>&1; tail -1 $OUT/smoke.txt
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.txt; tail -3 $OUT/pytest_gpu.txt
timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" >> $OUT/bench.err; tail -2 $OUT/bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/bench_reference.json 2>> $OUT/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2e/bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'frac',d['roofline']['frac'],'ms',d['ms_per_step'])
print('recon',d['reconstruct']['roofline_frac'],'sustained',{k:d['sustained'][k] for k in ('roofline_frac','ms_every_10th_step','low_power_variant_from_step','heat_ms_at_end')})
print('variant',d['variant_30000MiB']['roofline_frac'],'batch',{k:d['batch'][k] for k in ('roofline_frac','digest','rank0_launches_on_low_power_variant')})
print('e2e',d['e2e']['value'],'files',d['file_level'])
print('host_api',{n:{k:v for k,v in r.items() if k.endswith('encode_GBps') or k.startswith('cpu')} for n,r in d['host_api']['sizes'].items()})
r=json.loads(open('gpurun_out/r2e/bench_reference.json').read().strip().splitlines()[-1]); print('reference',r['value'],r['cpu_baseline']['kind'],r['cpu_baseline']['cores'])
PY
timeout 400 python scripts/bench_files_multi.py --gpus 1 --volumes 4 --gib 4 --rebuild > $OUT/files_multi.jsonl 2> $OUT/files_multi.err; cat $OUT/files_multi.jsonl | cut -c1-330; tail -2 $OUT/files_multi.err
SWEC_FILE_IO_PIECE=8388608 timeout 400 python scripts/bench_files_multi.py --gpus 1 --volumes 4 --gib 4 --no-cpu > $OUT/files_multi_piece8m.jsonl 2>> $OUT/files_multi.err; cat $OUT/files_multi_piece8m.jsonl | cut -c1-330
