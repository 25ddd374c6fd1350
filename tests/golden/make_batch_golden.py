#!/usr/bin/env python
"""tests/golden/make_batch_golden.py — expected checksum-of-checksums of BASELINE configs[3] (256 × 30 GiB synthetic
volumes, volume v seeded SEED0 + v), computed ENTIRELY on the CPU by the oracle (oracle/cpu_baseline.c
orc_volume_digests with the reference's own compiled C kernel when oracle/_ref is built): every volume is regenerated
from the seeded generator, walked through encodeDatFile's two-tier layout and encoded chunk by chunk; nothing of a
volume is held in memory.  bench.py's `batch` leg and tests/test_gpu_parity.py compare the device's result with the
committed numbers, so the 256-volume batch is oracle-checked byte for byte at every N without re-encoding 7.5 TiB on
the CPU of the GPU box (≈ 4 s per volume on 8 cores; this script ran for ≈ 20 minutes).

    python tests/golden/make_batch_golden.py [--volumes 256] [--threads N]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import pyoracle as po                      # noqa: E402
from seaweedfs_b200 import sharding                    # noqa: E402  (placement + combine rule only; no compute)

MASK = (1 << 64) - 1


def fold_parity(digests14):
    """Per-volume digest of bench.py's batch leg: the four parity-shard digests folded in shard order."""
    d = 0
    for one in digests14[10:]:
        d = (d * 0x100000001B3 + one) & MASK
    return d


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--volumes", type=int, default=256)
    ap.add_argument("--gib", type=float, default=30.0)
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--out", default=os.path.join(os.path.dirname(os.path.abspath(__file__)), "batch256.json"))
    a = ap.parse_args()
    dat_size = int(a.gib * (1 << 30))
    kind = po.best_cpu_kind()
    per_volume, shards = {}, {}
    t0 = time.time()
    for v in range(a.volumes):
        d14 = po.volume_digests(dat_size, sharding.volume_seed(v), threads=a.threads, kind=kind)
        shards[v] = ["%016x" % x for x in d14]
        per_volume[v] = fold_parity(d14)
        if v % 16 == 15:
            print(f"{v + 1}/{a.volumes} volumes, {time.time() - t0:.0f} s", flush=True)
    out = {"volumes": a.volumes, "dat_bytes_per_volume": dat_size, "seed0": hex(sharding.SEED0),
           "cpu_kind": {0: "reference C kernel (oracle/_ref)", 1: "GFNI port", 2: "scalar tables"}[kind],
           "digest": "%016x" % sharding.combine_digests(per_volume),
           "prefix_digests": {str(n): "%016x" % sharding.combine_digests({v: per_volume[v] for v in range(n)})
                              for n in (1, 2, 4, 8, 16, 32, 64, 128, 256) if n <= a.volumes},
           "per_volume": ["%016x" % per_volume[v] for v in range(a.volumes)],
           "shard_digests_volume0": shards[0]}
    with open(a.out, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", a.out, out["digest"])


if __name__ == "__main__":
    main()
