#!/usr/bin/env python
"""scripts/bench_files_multi.py — BASELINE configs[4] at its shape: `ec.encode` of several volumes at once through the
file-level entry points, volumes spread round-robin over the GPUs of ONE process (what a Go volume server does with
handle v on swecPickDevice(v); the shell runs up to 10 volumes concurrently: weed/shell/common.go:11,28-53,
command_ec_encode.go:302-315 → VolumeEcShardsGenerate, weed/server/volume_grpc_erasure_coding.go:43-146).

For every GPU count in --gpus: V = --volumes .dat files (distinct seeded synthetic volumes of --gib GiB in --dir) are
encoded concurrently, one thread per volume calling swec_write_ec_files(base_v, device = v mod G); aggregate .dat GB/s
= V·size / wall.  Then the CPU arm at the SAME schedule (oracle/cpu_baseline.c orc_generate_ec_files_mt: all host
threads, stripes in flight, reads ∥ GFNI ∥ writes) on the same volumes, and the reference-shaped serial walk.
Checks: every shard of every volume equals the CPU oracle's whole-volume digest for that seed (device digest of the
file bytes), i.e. byte-identical shards at every GPU count.  One JSON line per measurement."""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SEED0 = 0x5EA3EED5F00DCAFE


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dir", default="/dev/shm")
    ap.add_argument("--gib", type=float, default=8.0)
    ap.add_argument("--volumes", type=int, default=8)
    ap.add_argument("--gpus", default="1,2,4,8")
    ap.add_argument("--direct", type=int, default=0, help="file_direct_io option: 1 reads, 2 writes, 3 both")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--rebuild", action="store_true", help="also time ec.rebuild of 4 lost shards per volume")
    args = ap.parse_args()
    import numpy as np
    import torch
    import seaweedfs_b200
    from oracle import pyoracle as po
    from seaweedfs_b200 import erasure_coding as ec
    L = seaweedfs_b200.lib()
    ng = torch.cuda.device_count()
    d = os.path.join(args.dir, "swec_files_multi")
    os.makedirs(d, exist_ok=True)
    size = int(args.gib * (1 << 30)) + 12345
    bases = [os.path.join(d, str(100 + v)) for v in range(args.volumes)]
    assert L.swec_set_option(b"file_direct_io", args.direct) == 0

    def make(v):
        g = v % ng
        with torch.cuda.device(g):
            t = torch.empty((size + 7) & ~7, dtype=torch.uint8, device=f"cuda:{g}")
            assert L.swec_synth_fill_device(g, t.data_ptr(), 0, t.numel(), SEED0 + v, 0) == 0
            t[:size].cpu().numpy().tofile(bases[v] + ".dat")
    th = [threading.Thread(target=make, args=(v,)) for v in range(args.volumes)]
    [t.start() for t in th]
    [t.join() for t in th]
    torch.cuda.empty_cache()

    def drop_shards():
        for b in bases:
            for i in range(14):
                try:
                    os.remove(b + ec.ToExt(i))
                except FileNotFoundError:
                    pass

    def run_parallel(fn):
        errs = []
        bar = threading.Barrier(args.volumes + 1)

        def one(v):
            bar.wait()
            try:
                fn(v)
            except Exception as ex:                      # noqa: BLE001
                errs.append(repr(ex))
        th = [threading.Thread(target=one, args=(v,)) for v in range(args.volumes)]
        [t.start() for t in th]
        bar.wait()
        t0 = time.perf_counter()
        [t.join() for t in th]
        dt = time.perf_counter() - t0
        assert not errs, errs
        return dt

    def shard_digest(path, g=0):
        a = np.fromfile(path, dtype=np.uint8)
        with torch.cuda.device(g):
            t = torch.from_numpy(a).to(f"cuda:{g}")
            one = C.c_uint64(0)
            assert L.swec_digest_device(g, t.data_ptr(), t.numel(), C.byref(one), 0) == 0
        return one.value

    want = {v: po.volume_digests(size, SEED0 + v) for v in (0, args.volumes - 1)}
    common = {"volumes": args.volumes, "dat_bytes_per_volume": size, "dir": args.dir, "direct_io": args.direct}
    for G in [int(x) for x in args.gpus.split(",") if int(x) <= ng]:
        drop_shards()
        run_parallel(lambda v: ec.write_ec_files(bases[v], ec.NewDefaultECContext(device=v % G)))          # warm-up: rings, contexts
        drop_shards()
        dt = run_parallel(lambda v: ec.write_ec_files(bases[v], ec.NewDefaultECContext(device=v % G)))
        ok = all(shard_digest(bases[v] + ec.ToExt(i), v % G) == want[v][i] for v in want for i in range(14))
        row = dict(common, op="ec.encode (swec_write_ec_files)", gpus=G, seconds=round(dt, 3),
                   dat_GBps=round(args.volumes * size / dt / 1e9, 2), shards_equal_cpu_oracle=ok)
        if args.rebuild:
            for b in bases:
                for i in (1, 4, 10, 12):
                    os.remove(b + ec.ToExt(i))
            dt = run_parallel(lambda v: ec.rebuild_ec_files(bases[v], device=v % G))
            shard = os.path.getsize(bases[0] + ".ec00")
            ok2 = all(shard_digest(bases[v] + ec.ToExt(i), v % G) == want[v][i] for v in want for i in (1, 4, 10, 12))
            row.update(rebuild_seconds=round(dt, 3), rebuild_GBps_of_shard_bytes_read=round(args.volumes * 10 * shard / dt / 1e9, 2),
                       rebuilt_shards_equal_cpu_oracle=ok2)
        print(json.dumps(row), flush=True)
    if not args.no_cpu:
        threads = os.cpu_count() or 1
        try:
            q, per = open("/sys/fs/cgroup/cpu.max").read().split()
            if q != "max":
                threads = min(threads, max(1, int(int(q) / int(per))))
        except Exception:
            pass
        per_vol = max(1, threads // args.volumes)
        drop_shards()
        def cpu_one(v):
            if po.generate_ec_files_mt(bases[v], threads=per_vol) != 0:
                raise RuntimeError("cpu arm failed")
        dt = run_parallel(cpu_one)
        ok = all(shard_digest(bases[v] + ec.ToExt(i)) == want[v][i] for v in want for i in range(14))
        print(json.dumps(dict(common, op="CPU arm, same schedule (orc_generate_ec_files_mt, GFNI)", host_threads=per_vol * args.volumes,
                              seconds=round(dt, 3), dat_GBps=round(args.volumes * size / dt / 1e9, 2), shards_equal_cpu_oracle=ok)), flush=True)
        drop_shards()
        t0 = time.perf_counter()
        rc = po.generate_ec_files_simd(bases[0], 1 if po.gfni_level() else 0)
        dt = time.perf_counter() - t0
        print(json.dumps(dict(common, op="CPU arm, reference-shaped serial walk (one volume, one thread, 256 KiB batches)",
                              seconds=round(dt, 3), dat_GBps=round(size / dt / 1e9, 2), rc=rc)), flush=True)
    for f in os.listdir(d):
        os.remove(os.path.join(d, f))
    os.rmdir(d)


if __name__ == "__main__":
    main()
