#!/bin/bash
# round-2 session C: new tests, default bench line, batch probe, needles, file-level incl. O_DIRECT, ncu captures
OUT=gpurun_out/r2c; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.txt
tail -4 $OUT/pytest_gpu.txt
timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" >> $OUT/bench.err
tail -3 $OUT/bench.err; cat $OUT/bench.json
timeout 600 python scripts/experiments/batch_probe.py > $OUT/batch_probe.jsonl 2> $OUT/batch_probe.err; cat $OUT/batch_probe.jsonl; tail -3 $OUT/batch_probe.err
timeout 300 python scripts/bench_needles.py > $OUT/needles.jsonl 2> $OUT/needles.err; cat $OUT/needles.jsonl; tail -3 $OUT/needles.err
{ df -T /tmp /dev/shm /root 2>&1; lsblk 2>&1 | head -30; mount | grep -E " /tmp | / " ; } > $OUT/storage.txt 2>&1; cat $OUT/storage.txt
for dio in 0 1 3; do
  SWEC_FILE_DIRECT=$dio timeout 400 python scripts/bench_files.py --dir /tmp --gib 8 --cpu-gib 1 > $OUT/files_tmp_direct$dio.json 2>> $OUT/files.err
  echo "direct=$dio"; cat $OUT/files_tmp_direct$dio.json
done
# launch list of the default line (short), then full captures: low-power encode variant, AOT worst-case reconstruct, AOT single-loss
Q="--no-e2e --no-cpu-baseline --no-files --no-host-api --no-variant --no-sustained --batch-leg-volumes 8"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 $Q > $OUT/ncu_launch_run.txt 2>&1
SWEC_POWER_MODE=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:rs10x4_encode -s 3 -c 1 -o $OUT/prof_encode_lowpower \
    python bench.py --steps 2 --warmup 3 $Q --no-reconstruct --batch-leg-volumes 0 > $OUT/ncu_lp_run.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:swec_aot_recon -s 3 -c 1 -o $OUT/prof_aot_recon_worst \
    python bench.py --steps 2 --warmup 3 $Q --batch-leg-volumes 0 > $OUT/ncu_aot_run.txt 2>&1
ls -la $OUT
