#!/bin/bash
# round-2 session D (8 GPUs, one process group per measurement): what the host can feed to several GPUs, and the
# file-level path at BASELINE configs[4]'s shape.
OUT=gpurun_out/r2d; mkdir -p $OUT
nvidia-smi --query-gpu=index,pci.bus_id --format=csv > $OUT/gpus.txt 2>&1
for d in /sys/bus/pci/devices/*; do if [ "$(cat $d/class 2>/dev/null)" = "0x030200" ]; then echo "$d numa=$(cat $d/numa_node)"; fi; done >> $OUT/gpus.txt
cat /sys/kernel/mm/transparent_hugepage/enabled >> $OUT/gpus.txt 2>&1
# 1. plain DMA ceilings: 1/2/4 GPUs of one socket, one per socket, all 8 — huge pages vs 4 KiB pages
PROBE_GIB=2 timeout 300 python scripts/pcie_socket_probe.py > $OUT/pcie_probe_thp.jsonl 2> $OUT/pcie_probe.err
SWEC_NO_THP=1 PROBE_GIB=2 timeout 300 python scripts/pcie_socket_probe.py > $OUT/pcie_probe_4k.jsonl 2>> $OUT/pcie_probe.err
cat $OUT/pcie_probe_thp.jsonl; echo "--- 4k pages"; cat $OUT/pcie_probe_4k.jsonl
# 2. bench.py's e2e leg at N = 2, 4 (socket-interleaved placement vs LOCAL_RANK order) and N = 8
Q="--steps 5 --warmup 3 --no-cpu-baseline --no-files --no-host-api --no-variant --no-sustained --no-reconstruct --batch-leg-volumes 0 --e2e-gib 10"
run() { # n, tag, env...
  n=$1; tag=$2; shift 2
  env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n $Q > $OUT/bench_n${n}_$tag.json 2> $OUT/bench_n${n}_$tag.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("$OUT/bench_n${n}_$tag.json") if l.startswith("{")][-1])
    print("N=$n $tag", "value", d["value"], "e2e", d["e2e"]["value"], "e2e frac", d["e2e"]["roofline"]["frac"], d["config"].get("rank0_device"))
except Exception as ex:
    print("N=$n $tag FAILED", ex)
PY
}
run 4 spread SWEC_X=1
run 4 nospread SWEC_BENCH_NO_SPREAD=1
run 2 spread SWEC_X=1
run 2 nospread SWEC_BENCH_NO_SPREAD=1
run 8 all SWEC_X=1
# 3. ONE Encoder.Encode call split by column range over the GPUs, host buffers laid out per GPU's NUMA node
timeout 400 python scripts/bench_group.py --gib 20 --numa-split > $OUT/group_numa_split.jsonl 2> $OUT/group.err; cat $OUT/group_numa_split.jsonl
timeout 300 python scripts/bench_group.py --gib 20 > $OUT/group_one_node.jsonl 2>> $OUT/group.err; cat $OUT/group_one_node.jsonl
# 4. configs[4] at shape: 8 volumes x 8 GiB through the file-level entry points over 1/2/4/8 GPUs of one process, tmpfs
timeout 900 python scripts/bench_files_multi.py --gpus 1,2,4,8 --volumes 8 --gib 8 --rebuild > $OUT/files_multi_shm.jsonl 2> $OUT/files_multi.err; cat $OUT/files_multi_shm.jsonl; tail -3 $OUT/files_multi.err
# 5. ... and on the overlay (NVMe-backed) file system with O_DIRECT reads, 4 volumes x 4 GiB
timeout 600 python scripts/bench_files_multi.py --gpus 1,4 --volumes 4 --gib 4 --dir /tmp --direct 1 --no-cpu > $OUT/files_multi_tmp_direct1.jsonl 2>> $OUT/files_multi.err; cat $OUT/files_multi_tmp_direct1.jsonl
ls $OUT
