#!/usr/bin/env python
"""scripts/bench_files.py — BASELINE configs[4] in miniature: file-level WriteEcFiles / RebuildEcFiles
through the C ABI (pread → pinned slot → H2D → kernel → D2H → pwrite, three slots in flight) next to
the reference-shaped CPU walk (oracle generate_ec_files: serial 256 KiB read → Encode → write, one
thread, like ec_encoder.go:248-278).  Files live in --dir (default /dev/shm: RAM-backed, so the
number is the pipeline's, not a disk's)."""
import argparse
import hashlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        while True:
            b = f.read(1 << 24)
            if not b:
                break
            h.update(b)
    return h.hexdigest()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dir", default="/dev/shm")
    ap.add_argument("--gib", type=float, default=4.0)
    ap.add_argument("--cpu-gib", type=float, default=1.0)
    print(json.dumps(run(ap.parse_args())))


def run(args):
    """args: .dir, .gib, .cpu_gib → result dict (also called by bench.py's file-level leg)"""
    import numpy as np
    import torch
    import seaweedfs_b200
    from oracle import pyoracle as po
    from seaweedfs_b200 import erasure_coding as ec
    L = seaweedfs_b200.lib()
    d = os.path.join(args.dir, "swec_bench_files")
    os.makedirs(d, exist_ok=True)
    base = os.path.join(d, "7")
    size = int(args.gib * (1 << 30)) + 12345
    # synthetic .dat from the device generator
    t = torch.empty((size + 7) & ~7, dtype=torch.uint8, device="cuda")
    L.swec_synth_fill_device(0, t.data_ptr(), 0, t.numel(), 0x5EA3EED5F00DCAFE, 0)
    t[:size].cpu().numpy().tofile(base + ".dat")
    del t
    out = {"dat_bytes": size, "dir": args.dir}
    ec.write_ec_files(base)                                       # warm-up (context, page cache)
    t0 = time.perf_counter()
    ec.write_ec_files(base)                                       # over the warm-up's shard files: O_TRUNC frees 1.4x the
    dt = time.perf_counter() - t0                                 # volume in page-cache / tmpfs pages before the first byte
    out["write_ec_files_over_existing_shards_GBps"] = round(size / dt / 1e9, 2)
    for i in range(14):
        os.remove(base + ec.ToExt(i))
    t0 = time.perf_counter()
    ec.write_ec_files(base)                                       # what ec.encode does: the shard files do not exist yet
    dt = time.perf_counter() - t0
    out["write_ec_files_GBps"] = round(size / dt / 1e9, 2)
    digests = [sha(base + ec.ToExt(i)) for i in range(14)]
    for i in (1, 4, 10, 12):
        os.remove(base + ec.ToExt(i))
    t0 = time.perf_counter()
    rebuilt = ec.rebuild_ec_files(base)
    dt = time.perf_counter() - t0
    out["rebuild_4_shards_GBps_of_shard_bytes_read"] = round(10 * os.path.getsize(base + ".ec00") / dt / 1e9, 2)
    assert rebuilt == [1, 4, 10, 12] and [sha(base + ec.ToExt(i)) for i in range(14)] == digests
    # CPU walk shaped like the reference (single goroutine, 256 KiB batches) on a smaller file
    csize = int(args.cpu_gib * (1 << 30)) + 12345
    cbase = os.path.join(d, "8")
    with open(base + ".dat", "rb") as f, open(cbase + ".dat", "wb") as g:
        g.write(f.read(csize))
    walks = {}
    for kind, name in ((0, "reference_c_kernel"), (1, "gfni_port")):
        t0 = time.perf_counter()
        rc = po.generate_ec_files_simd(cbase, kind)
        if rc == 0:
            walks[name] = round(csize / (time.perf_counter() - t0) / 1e9, 3)
    out["cpu_reference_shaped_walk_GBps"] = walks
    out["cpu_walk_note"] = ("generateEcFiles as the reference schedules it: one thread, 256 KiB batches, "
                            "pread x10 -> Encode (SIMD) -> write x14, strictly serial")
    if not walks:
        assert po.generate_ec_files(cbase) == 0
    # same bytes from both paths on the common prefix? (different sizes → compare via oracle on GPU output instead)
    gp = os.path.join(d, "9")
    os.link(cbase + ".dat", gp + ".dat")
    ec.write_ec_files(gp)
    out["gpu_files_equal_cpu_files"] = all(sha(gp + ec.ToExt(i)) == sha(cbase + ec.ToExt(i)) for i in range(14))
    for f in os.listdir(d):
        os.remove(os.path.join(d, f))
    os.rmdir(d)
    return out


if __name__ == "__main__":
    main()
