O=gpurun_out/r2s; mkdir -p $O
for dir in /dev/shm /tmp; do
  python scripts/bench_decode.py $dir 8 >> $O/decode.txt 2>&1
  SWEC_IO_THREADS=1 python scripts/bench_decode.py $dir 8 2>&1 | sed 's/^/io_threads=1 /' >> $O/decode.txt
done
cat $O/decode.txt
timeout 600 python -m pytest tests/test_volume_ops.py tests/test_gpu_reference_suites.py -m gpu -x -q 2>&1 | tail -2
