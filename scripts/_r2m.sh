O=gpurun_out/r2m; mkdir -p $O
gcc -O2 -o /tmp/io_probe scripts/experiments/io_probe.c -lpthread || exit 1
df -T /tmp /dev/shm > $O/fs.txt 2>&1; cat /sys/kernel/mm/transparent_hugepage/enabled >> $O/fs.txt 2>&1
for dir in /dev/shm /tmp; do
  for cfg in "16 3 0 1" "16 3 0 0" "16 1 0 1" "16 5 0 1" "32 5 0 1" "4 3 0 1" "16 3 1 1"; do
    set -- $cfg
    timeout 120 /tmp/io_probe $dir 8 $1 $2 $3 $4 >> $O/io_probe.jsonl 2>> $O/io.err
  done
done
cat $O/fs.txt; cat $O/io_probe.jsonl; tail -3 $O/io.err
