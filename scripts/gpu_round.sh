#!/bin/bash
# scripts/gpu_round.sh — one gpurun session: probes, GPU tests, smoke, bench, sweeps, launch list, ncu capture.
# usage (from the repo root on the GPU box): bash scripts/gpu_round.sh [tag] [sections]
TAG=${1:-r1}
SECTIONS=${2:-"env smoke tests bench sweep e2e ncu"}
OUT=gpurun_out/$TAG
mkdir -p $OUT
has() { [[ " $SECTIONS " == *" $1 "* ]]; }
QUICK="--no-e2e --no-cpu-baseline --no-reconstruct"
if has env; then
{
  echo "== host"; nproc; grep -m1 'model name' /proc/cpuinfo; free -g | head -2; lscpu | grep -i numa
  grep -m1 flags /proc/cpuinfo | tr ' ' '\n' | grep -E 'avx512f|avx512bw|gfni|avx2|ssse3' | tr '\n' ' '; echo
  echo "== gpu"; nvidia-smi --query-gpu=index,name,memory.total,clocks.max.sm,clocks.max.mem,power.limit,pci.bus_id --format=csv
  for d in /sys/bus/pci/devices/*; do if [ "$(cat $d/class 2>/dev/null)" = "0x030200" ]; then echo "$d numa=$(cat $d/numa_node)"; fi; done
  nvidia-smi topo -m
} > $OUT/env.txt 2>&1
fi
if has smoke; then python __graft_entry__.py --smoke > $OUT/smoke.txt 2>&1; echo "smoke rc=$?" >> $OUT/smoke.txt; fi
if has tests; then timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.txt; fi
if has bench; then
  timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" >> $OUT/bench.err
  timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/bench_reference.json 2>> $OUT/bench.err
fi
if has sweep; then
  for shape in "128 1" "128 2" "256 1" "256 2" "512 1" "512 2"; do
    set -- $shape
    SWEC_ENC_THREADS=$1 SWEC_ENC_UNROLL=$2 timeout 200 python bench.py --steps 5 --warmup 3 $QUICK 2>&1 | sed "s/^/threads=$1 unroll=$2 /" >> $OUT/sweep.txt
  done
  for c in 2 3 5 6 8; do
    SWEC_CTAS_PER_SM=$c timeout 200 python bench.py --steps 5 --warmup 3 $QUICK 2>&1 | sed "s/^/threads=256 unroll=1 ctas=$c /" >> $OUT/sweep.txt
  done
fi
if has e2e; then
  for cfg in "4194304 3 0" "16777216 3 0" "16777216 4 0" "33554432 4 0" "16777216 4 1"; do
    set -- $cfg
    if [ "$3" = "1" ]; then export SWEC_NO_NUMA=1; else unset SWEC_NO_NUMA; fi
    SWEC_STAGE_CHUNK=$1 SWEC_STAGE_SLOTS=$2 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-reconstruct --e2e-gib 10 2>&1 \
      | python -c "import sys,json; [print('chunk=$1 slots=$2 no_numa=$3', json.loads(l)['e2e']) for l in sys.stdin if l.startswith('{')]" >> $OUT/e2e_sweep.txt
  done
  unset SWEC_NO_NUMA
fi
if has ncu; then
  # every launch with its device time (cold-cache, serialised: compare shares)
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches.csv \
      python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/ncu_launch_run.txt 2>&1
  # the top kernel, full set, one launch over the whole 30 GiB volume
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:rs10x4_encode -s 3 -c 1 -o $OUT/prof_encode \
      python bench.py --steps 2 --warmup 3 $QUICK > $OUT/ncu_full_run.txt 2>&1
  # the specialised reconstruct kernel
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:swec_jit -s 3 -c 1 -o $OUT/prof_reconstruct \
      python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/ncu_full_recon_run.txt 2>&1
fi
for f in smoke.txt pytest_gpu.txt bench.err; do [ -f $OUT/$f ] && tail -n 3 $OUT/$f; done
for f in bench.json bench_reference.json sweep.txt e2e_sweep.txt; do [ -f $OUT/$f ] && cat $OUT/$f; done
exit 0
