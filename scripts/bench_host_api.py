#!/usr/bin/env python
"""scripts/bench_host_api.py — the Encoder-level C ABI on PAGEABLE host buffers at the batch sizes the
Go code uses (256 KiB per shard in encodeDataOneBatch, 1 MiB in rebuildEcFiles, needle-sized degraded
reads), next to one CPU thread of the reference arithmetic on the same buffers."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def rate(fn, min_s=0.5):
    fn()
    t0, n = time.perf_counter(), 0
    while time.perf_counter() - t0 < min_s:
        fn()
        n += 1
    return n / (time.perf_counter() - t0)


def main():
    import seaweedfs_b200
    from oracle import pyoracle as po
    from seaweedfs_b200 import erasure_coding as ec
    enc = ec.Encoder(10, 4, device=0)
    rows = po.build_matrix(10, 14)[10:]
    rng = np.random.default_rng(0)
    for n in (4096, 65536, 256 * 1024, 1 << 20, 16 << 20):
        shards = [rng.integers(0, 256, n, dtype=np.uint8) for _ in range(10)] + [np.zeros(n, dtype=np.uint8) for _ in range(4)]
        r_enc = rate(lambda: enc.encode(shards))
        holes = list(shards)

        def recon():
            holes[5] = None
            enc.reconstruct_data(holes)
        r_rec = rate(recon)
        kind = 1 if po.gfni_level() else 0
        outs = [np.zeros(n, dtype=np.uint8) for _ in range(4)]
        r_cpu = rate(lambda: po.cpu_apply(kind, rows, shards[:10], outs, threads=1))
        print(json.dumps({"shard_bytes": n, "encode_calls_per_s": round(r_enc, 1),
                          "encode_input_GBps": round(r_enc * 10 * n / 1e9, 3),
                          "reconstruct_data_calls_per_s": round(r_rec, 1),
                          "reconstruct_input_GBps": round(r_rec * 10 * n / 1e9, 3),
                          "cpu_1thread_input_GBps": round(r_cpu * 10 * n / 1e9, 3)}), flush=True)


if __name__ == "__main__":
    main()
