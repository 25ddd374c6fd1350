#!/bin/bash
# round-2 session B: new GPU tests, host-API sweep (DMA ring vs zero-copy, pieces, copy crew), power-mode legs
OUT=gpurun_out/r2b; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.txt
tail -5 $OUT/pytest_gpu.txt
timeout 900 python -m pytest tests/test_gpu_jit_cache.py -m gpu -x -q -s > $OUT/jit_cache.txt 2>&1
grep first_process $OUT/jit_cache.txt
timeout 600 python scripts/bench_host_api.py > $OUT/host_api_sweep.jsonl 2> $OUT/host_api_sweep.err
SWEC_HOST_COPY_THREADS=1 timeout 300 python scripts/bench_host_api.py --quick > $OUT/host_api_sweep_1thread.jsonl 2>> $OUT/host_api_sweep.err
SWEC_HOST_COPY_SPIN_US=0 timeout 300 python scripts/bench_host_api.py --quick > $OUT/host_api_sweep_nospin.jsonl 2>> $OUT/host_api_sweep.err
cat $OUT/host_api_sweep.jsonl; echo ---1thread; cat $OUT/host_api_sweep_1thread.jsonl; echo ---nospin; cat $OUT/host_api_sweep_nospin.jsonl
Q="--no-e2e --no-cpu-baseline --no-files --no-host-api --no-variant"
for pm in 1 2 0; do
  SWEC_POWER_MODE=$pm timeout 300 python bench.py --steps 10 --warmup 3 $Q 2>>$OUT/power.err | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(json.dumps({'power_mode':$pm,'burst_ms':d['ms_per_step'],'frac':d['roofline']['frac'],'recon':d['reconstruct'],'sustained':d['sustained'],'batch':{k:d['batch'][k] for k in ('ms_per_volume','roofline_frac','digest','sm_mhz','power_w_max')}}))" >> $OUT/power_modes.jsonl
done
cat $OUT/power_modes.jsonl
