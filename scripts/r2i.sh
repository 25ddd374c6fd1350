#!/bin/bash
# round-2 session I (8 GPUs, short): configs[4] at shape again, now with the process-wide I/O pool
OUT=gpurun_out/r2i; mkdir -p $OUT
timeout 500 python scripts/bench_files_multi.py --gpus 1,8 --volumes 8 --gib 8 --rebuild > $OUT/files_multi_shm.jsonl 2> $OUT/files_multi.err; cut -c1-330 $OUT/files_multi_shm.jsonl; tail -2 $OUT/files_multi.err
