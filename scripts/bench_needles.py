#!/usr/bin/env python
"""scripts/bench_needles.py — degraded reads at file level: every needle of a synthetic EC volume read through
swec_read_ec_needles with four shard files missing, (a) all needles in ONE call (one batched ReconstructData),
(b) one call per needle (what store_ec.go:482-560 does today: one reedsolomon.New + ReconstructData each).
Results are checked against the volume image."""
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


CAP = 1 << 17     # per-needle buffer; small on purpose: 1 MiB strides make every first touch a 2 MiB THP fault


def main():
    import seaweedfs_b200
    from oracle import pyoracle as po
    from oracle import rs_numpy as rn
    from seaweedfs_b200 import erasure_coding as ec
    from test_volume_ops import expected_record, synthetic_volume
    L = seaweedfs_b200.lib()
    dat, idx = synthetic_volume(seed=5, needles=4000)
    with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as d:
        base = os.path.join(d, "9")
        for i, s in enumerate(po.encode_dat_image(dat)):
            s.tofile(base + ".ec%02d" % i)
        open(base + ".ecx", "wb").write(rn.sorted_ecx_from_idx(idx))
        json.dump({"version": 3, "datFileSize": str(len(dat)), "ecShardConfig": {"dataShards": 10, "parityShards": 4}},
                  open(base + ".vif", "w"))
        live = list(rn._entries(rn.sorted_ecx_from_idx(idx)))
        ids = [k for k, _, _ in live]
        for lost in ((), (0, 1, 2, 3)):
            for i in lost:
                os.remove(base + ".ec%02d" % i)
            ec.ReadEcShardNeedles(base, ids[:8], capacity=CAP)           # warm-up (tables, staging ring)
            l0, t0 = L.swec_kernel_launches(), time.perf_counter()
            out = ec.ReadEcShardNeedles(base, ids, capacity=CAP)
            t_batch, l_batch = time.perf_counter() - t0, L.swec_kernel_launches() - l0
            for (key, off, size), r in zip(live, out):
                assert r["status"] == "SWEC_OK" and (r["bytes"] == expected_record(dat, off * 8, size)).all()
            sample = ids[:400]
            l0, t0 = L.swec_kernel_launches(), time.perf_counter()
            for nid in sample:
                ec.ReadEcShardNeedles(base, [nid], capacity=CAP)
            t_single, l_single = (time.perf_counter() - t0) / len(sample), (L.swec_kernel_launches() - l0) / len(sample)
            vol = ec.EcVolume(base)
            vol.ReadEcShardNeedles(ids)                                  # grows the staging ring once
            t0 = time.perf_counter()
            out2 = vol.ReadEcShardNeedles(ids)                           # steady state, exact-size arena (two passes)
            t_mounted_batch = time.perf_counter() - t0
            assert all((a["bytes"] == b["bytes"]).all() for a, b in zip(out, out2))
            t0 = time.perf_counter()
            for nid in sample:
                vol.ReadEcShardNeedles([nid], capacity=CAP)
            t_mounted_single = (time.perf_counter() - t0) / len(sample)
            vol.close()
            print(json.dumps({"volume_MiB": round(len(dat) / 2**20, 1), "needles": len(ids), "lost_shards": list(lost),
                              "recovered_intervals": sum(r["recovered_intervals"] for r in out),
                              "one_call_all_needles": {"needles_per_s": round(len(ids) / t_batch), "MBps": round(sum(r["n_bytes"] for r in out) / t_batch / 1e6, 1), "gpu_launches": int(l_batch)},
                              "one_call_per_needle": {"needles_per_s": round(1 / t_single), "gpu_launches_per_needle": round(l_single, 2)},
                              "mounted_volume": {"all_needles_one_call_per_s": round(len(ids) / t_mounted_batch),
                                                 "one_needle_per_call_per_s": round(1 / t_mounted_single)}}),
                  flush=True)


if __name__ == "__main__":
    main()
