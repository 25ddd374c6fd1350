O=gpurun_out/r2n; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "rc=$?" >> $O/pytest_gpu.txt
for dir in /dev/shm /tmp; do
  SWEC_PIPE_STATS=1 timeout 600 python scripts/bench_files.py --dir $dir --gib 8 --cpu-gib 1 >> $O/files.txt 2>> $O/files_stats.txt; echo "rc=$? dir=$dir" >> $O/files.txt
done
grep -E "passed|failed|rc=" $O/pytest_gpu.txt | tail -2; cat $O/files.txt | cut -c1-600; grep generate_ec_files $O/files_stats.txt | head -8
