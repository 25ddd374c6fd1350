#!/usr/bin/env python
"""scripts/bench_group.py — ONE process, ONE Encoder.Encode call on pinned host shards, split by byte-column
range over 1, 2, 4, … GPUs of the box (swec_encode_multi).  Prints input GB/s per GPU count; the 1-GPU line
is bench.py's e2e leg.  Result checked against the single-handle call."""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GIB = 1 << 30


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gib", type=float, default=20.0, help="volume size (GiB of input)")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--interleave", action="store_true", help="plain cudaHostAlloc memory instead of NUMA-bound")
    ap.add_argument("--numa-split", action="store_true",
                    help="buffers from swec_alloc_pinned_shards: range g of every shard on the NUMA node of GPU g (re-laid per GPU count)")
    args = ap.parse_args()
    import torch
    import seaweedfs_b200
    from seaweedfs_b200 import erasure_coding as ec
    L = seaweedfs_b200.lib()
    ngpu = torch.cuda.device_count()
    n = int(args.gib * GIB / 10) & ~4095
    raw = L.swec_alloc_pinned(14 * n) if args.interleave else L.swec_alloc_pinned_for_device(0, 14 * n)
    assert raw
    host = np.ctypeslib.as_array(C.cast(raw, C.POINTER(C.c_uint8)), shape=(14 * n,))
    rng = np.random.default_rng(1)
    block = rng.integers(0, 256, 64 << 20, dtype=np.uint8)
    for off in range(0, 10 * n, len(block)):
        m = min(len(block), 10 * n - off)
        host[off:off + m] = block[:m]
        block = np.roll(block, 7919)
    shards = (C.c_void_p * 14)(*[raw + i * n for i in range(14)])
    ref_digest = None
    counts = [c for c in (1, 2, 4, 8) if c <= ngpu]
    for cnt in counts:
        grp = ec.EncoderGroup(10, 4, list(range(cnt)))
        split_base = None
        if args.numa_split:
            ptrs = (C.c_void_p * 14)()
            assert L.swec_alloc_pinned_shards(grp._arr, cnt, 14, n, ptrs) == 0
            split_base = ptrs[0]
            for i in range(10):                                   # same data as the plain buffer
                C.memmove(ptrs[i], raw + i * n, n)
            shards = ptrs
        for _ in range(2):
            assert L.swec_encode_multi(grp._arr, cnt, shards, n) == 0
        t0 = time.perf_counter()
        for _ in range(args.steps):
            assert L.swec_encode_multi(grp._arr, cnt, shards, n) == 0
        dt = (time.perf_counter() - t0) / args.steps
        par = host[10 * n:] if split_base is None else np.concatenate(
            [np.ctypeslib.as_array(C.cast(shards[10 + p], C.POINTER(C.c_uint8)), shape=(n,)) for p in range(4)])
        digest = int(np.bitwise_xor.reduce(par.view(np.uint64)))
        if ref_digest is None:
            ref_digest = digest
        print(json.dumps({"gpus_in_one_call": cnt, "input_GiB": round(10 * n / GIB, 2), "seconds": round(dt, 4),
                          "input_GBps": round(10 * n / dt / 1e9, 2), "parity_equals_1gpu_result": digest == ref_digest,
                          "host_memory": ("range g on the node of GPU g (swec_alloc_pinned_shards)" if args.numa_split else
                                          "cudaHostAlloc" if args.interleave else "NUMA node of GPU 0")}), flush=True)
        if split_base is not None:
            L.swec_free_pinned(split_base)
            shards = (C.c_void_p * 14)(*[raw + i * n for i in range(14)])
        grp.close()
    L.swec_free_pinned(raw)


if __name__ == "__main__":
    main()
