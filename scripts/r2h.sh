#!/bin/bash
# round-2 session H (1 GPU): degraded-read rates after limiting the packed zero-copy path to needle-sized batches
OUT=gpurun_out/r2h; mkdir -p $OUT
timeout 200 python scripts/bench_needles.py > $OUT/needles.jsonl 2> $OUT/needles.err; cat $OUT/needles.jsonl
SWEC_HOST_ZERO_COPY=0 timeout 200 python scripts/bench_needles.py > $OUT/needles_no_zero_copy.jsonl 2>> $OUT/needles.err; cat $OUT/needles_no_zero_copy.jsonl
timeout 300 python -m pytest tests/test_volume_ops.py tests/test_gpu_parity.py -m gpu -x -q -k "needle or degraded or batch or seam" > $OUT/pytest_sub.txt 2>&1; tail -2 $OUT/pytest_sub.txt
