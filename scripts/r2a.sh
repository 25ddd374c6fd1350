#!/bin/bash
# round-2 session A: tests, default bench line with the new legs, host-API sizes
OUT=gpurun_out/r2a; mkdir -p $OUT
bash scripts/gpu_round.sh r2a "env smoke tests" > $OUT/round.txt 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" >> $OUT/bench.err
tail -3 $OUT/pytest_gpu.txt; tail -5 $OUT/bench.err; cat $OUT/bench.json
