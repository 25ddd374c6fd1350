O=gpurun_out/r2r; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "rc=$?" >> $O/pytest_gpu.txt
for dir in /dev/shm /tmp; do
  SWEC_PIPE_STATS=1 timeout 600 python scripts/bench_files.py --dir $dir --gib 8 --cpu-gib 1 >> $O/files.txt 2>> $O/files_stats.txt; echo "rc=$? dir=$dir" >> $O/files.txt
done
grep -E "passed|failed|rc=" $O/pytest_gpu.txt | tail -2
python - <<'PY'
import json
for l in open('gpurun_out/r2r/files.txt'):
    if l.startswith('{'):
        d=json.loads(l); print(d['dir'], 'fresh', d['write_ec_files_GBps'], 'over-existing', d['write_ec_files_over_existing_shards_GBps'], 'rebuild', d['rebuild_4_shards_GBps_of_shard_bytes_read'], 'cpu walk', d['cpu_reference_shaped_walk_GBps'], d['gpu_files_equal_cpu_files'])
PY
grep '"pipe"' gpurun_out/r2r/files_stats.txt | cut -c1-330 | head -12
