#!/bin/bash
# scripts/gpu_round.sh — one gpurun session: probes, GPU tests, smoke, bench, launch list, ncu capture.
# usage (from the repo root on the GPU box): bash scripts/gpu_round.sh [tag]
TAG=${1:-r1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
{
  echo "== host"; nproc; grep -m1 'model name' /proc/cpuinfo; free -g | head -2
  grep -m1 flags /proc/cpuinfo | tr ' ' '\n' | grep -E 'avx512f|avx512bw|gfni|avx2|ssse3' | tr '\n' ' '; echo
  echo "== gpu"; nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem,power.limit --format=csv
  ls /usr/local/cuda/lib64/libnvrtc.so* 2>/dev/null | head -3
} > $OUT/env.txt 2>&1
python __graft_entry__.py --smoke > $OUT/smoke.txt 2>&1; echo "smoke rc=$?" >> $OUT/smoke.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.txt
timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" >> $OUT/bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2>> $OUT/bench.err
for c in 2 3 5 6 8; do
  SWEC_CTAS_PER_SM=$c timeout 200 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | sed "s/^/ctas=$c /" >> $OUT/sweep.txt
done
# every launch with its device time
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --volume-gib 10 > $OUT/ncu_launch_run.txt 2>&1
# the top kernel, full set
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rs10x4_encode -s 3 -c 1 -o $OUT/prof_encode \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --volume-gib 10 > $OUT/ncu_full_run.txt 2>&1
tail -3 $OUT/smoke.txt $OUT/pytest_gpu.txt $OUT/bench.err
cat $OUT/bench.json $OUT/sweep.txt
