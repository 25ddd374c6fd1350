#!/bin/bash
# round-2 session C2: validation of the multi-GPU scripts on one GPU, power-policy probe, fixed tests
OUT=gpurun_out/r2c2; mkdir -p $OUT
timeout 120 python scripts/experiments/power_policy_probe.py > $OUT/power_policy.txt 2>&1; head -40 $OUT/power_policy.txt
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 12 python scripts/experiments/power_policy_probe.py > $OUT/power_policy_ncu.txt 2>&1; grep -E "swec\]|==" $OUT/power_policy_ncu.txt | head -20
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.txt; tail -4 $OUT/pytest_gpu.txt
PROBE_GIB=2 timeout 200 python scripts/pcie_socket_probe.py > $OUT/pcie_probe.jsonl 2> $OUT/pcie_probe.err; cat $OUT/pcie_probe.jsonl; tail -3 $OUT/pcie_probe.err
timeout 400 python scripts/bench_files_multi.py --gpus 1 --volumes 2 --gib 2 --rebuild > $OUT/files_multi.jsonl 2> $OUT/files_multi.err; cat $OUT/files_multi.jsonl; tail -5 $OUT/files_multi.err
timeout 300 python scripts/bench_files.py --dir /dev/shm --gib 8 --cpu-gib 1 > $OUT/files_shm.json 2>> $OUT/files_multi.err; cat $OUT/files_shm.json
