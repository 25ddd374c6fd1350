#!/bin/bash
# round-2 session F (1 GPU): the process-wide I/O pool and read-piece size at file level; fresh launch list
OUT=gpurun_out/r2f; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.txt; tail -3 $OUT/pytest_gpu.txt
for piece in 2097152 8388608 2097152 8388608; do
  SWEC_FILE_IO_PIECE=$piece timeout 300 python scripts/bench_files.py --dir /dev/shm --gib 8 --cpu-gib 0.25 2>>$OUT/files.err | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(json.dumps({'read_piece':$piece, **{k:d[k] for k in ('write_ec_files_GBps','write_ec_files_over_existing_shards_GBps','rebuild_4_shards_GBps_of_shard_bytes_read')}}))" | tee -a $OUT/files_piece_sweep.jsonl
done
timeout 400 python scripts/bench_files_multi.py --gpus 1 --volumes 4 --gib 4 --rebuild > $OUT/files_multi_shared_pool.jsonl 2>> $OUT/files.err; cut -c1-330 $OUT/files_multi_shared_pool.jsonl
SWEC_IO_THREADS=32 timeout 400 python scripts/bench_files_multi.py --gpus 1 --volumes 4 --gib 4 --no-cpu > $OUT/files_multi_shared_pool_32.jsonl 2>> $OUT/files.err; cut -c1-330 $OUT/files_multi_shared_pool_32.jsonl
tail -3 $OUT/files.err
Q="--no-e2e --no-cpu-baseline --no-files --no-host-api --no-sustained --batch-leg-volumes 8"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 $Q > $OUT/ncu_launch_run.txt 2>&1; grep -c rs10x4 $OUT/launches.csv
