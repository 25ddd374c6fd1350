#!/usr/bin/env python
"""scripts/tune_encode.py — in-process sweep of the Horner kernel launch shape on one GPU.
One 30 GiB volume is generated once; every (threads, unroll, ctas_per_sm) shape is timed with CUDA
events over the same volume.  Prints one JSON line per shape (input GB/s, roofline fraction)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--volume-gib", type=float, default=30.0)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--reconstruct", action="store_true")
    ap.add_argument("--best", action="store_true", help="only the leading shapes, three repeats each")
    ap.add_argument("--tag", default="")
    ap.add_argument("--memprobe", action="store_true",
                    help="memory-pipeline ceiling: the same kernel body with a trivial (all-ones) matrix, and a D2D copy")
    args = ap.parse_args()
    import torch
    import seaweedfs_b200
    from seaweedfs_b200 import erasure_coding as ec
    L = seaweedfs_b200.lib()
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    enc = ec.Encoder(10, 4, device=0)
    stream = torch.cuda.current_stream().cuda_stream
    size = int(args.volume_gib * (1 << 30))
    shard = ec.expected_shard_size(size)
    dat = torch.empty(size, dtype=torch.uint8, device="cuda")
    par = [torch.empty(shard, dtype=torch.uint8, device="cuda") for _ in range(4)]
    L.swec_synth_fill_device(0, dat.data_ptr(), 0, size, 0x5EA3EED5F00DCAFE, stream)
    pp = [p.data_ptr() for p in par]

    def timed(fn):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.steps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / args.steps

    shapes = []
    for t, cs in ((128, (4, 5, 6, 8)), (256, (2, 3, 4)), (512, (1, 2))):
        for c in cs:
            shapes.append((t, 1, c))
    for t, cs in ((128, (2, 3, 4)), (256, (1, 2)), (512, (1,))):
        for c in cs:
            shapes.append((t, 2, c))
    if args.best:
        shapes = [(512, 2, 1), (128, 1, 5), (256, 1, 3), (128, 2, 3), (128, 1, 6), (256, 2, 2), (128, 2, 4)] * 3
    for t, u, c in shapes:
        assert L.swec_set_option(b"enc_threads", t) == 0
        assert L.swec_set_option(b"enc_unroll", u) == 0
        assert L.swec_set_option(b"ctas_per_sm", c) == 0
        ms = timed(lambda: enc.encode_volume_device(dat.data_ptr(), size, pp, stream))
        print(json.dumps({"tag": args.tag, "kernel": "rs10x4_encode", "threads": t, "unroll": u, "ctas_per_sm": c,
                          "ms": round(ms, 4), "input_GBps": round(size / ms / 1e6, 1),
                          "frac": round(1.4 * size / ms / 1e6 / peak, 4)}), flush=True)
    if args.memprobe:
        import numpy as np
        S = shard & ~15
        d = [dat.data_ptr() + i * S for i in range(10)]
        L.swec_set_option(b"jit_min_bytes", 1)
        for t, u, c in ((512, 2, 1), (256, 1, 3)):
            L.swec_set_option(b"enc_threads", t)
            L.swec_set_option(b"enc_unroll", u)
            L.swec_set_option(b"ctas_per_sm", c)
            e3 = ec.Encoder(10, 4, device=0)
            ones = np.ones((4, 10), dtype=np.uint8)
            ms = timed(lambda: e3.apply_device(ones, d, pp, S, stream))
            print(json.dumps({"kernel": "xor_all_10_to_4 (no GF work)", "threads": t, "unroll": u, "ctas_per_sm": c,
                              "ms": round(ms, 4), "frac": round(14 * S / ms / 1e6 / peak, 4)}), flush=True)
        a = dat[: 12 << 30]
        b = torch.empty(12 << 30, dtype=torch.uint8, device="cuda")
        ms = timed(lambda: b.copy_(a))
        print(json.dumps({"kernel": "torch copy 12 GiB", "ms": round(ms, 4), "GBps_rw": round(2 * (12 << 30) / ms / 1e6, 1)}), flush=True)
        del b
        L.swec_set_option(b"ctas_per_sm", 0)
    if args.reconstruct:
        S = shard & ~15
        d = [dat.data_ptr() + i * S for i in range(10)]
        scratch = [torch.empty(S, dtype=torch.uint8, device="cuda") for _ in range(4)]
        ptrs = [t.data_ptr() for t in scratch] + d[4:] + pp
        present = [0] * 4 + [1] * 10
        L.swec_set_option(b"enc_threads", 256)
        L.swec_set_option(b"enc_unroll", 1)
        for label, jit_min, cs in (("swec_jit", 1, (2, 3, 4)), ("swec_table_kernel", 1 << 60, (0,))):
            L.swec_set_option(b"jit_min_bytes", jit_min)
            e2 = ec.Encoder(10, 4, device=0)
            for c in cs:
                L.swec_set_option(b"ctas_per_sm", c)
                ms = timed(lambda: e2.reconstruct_device(ptrs, present, S, False, stream))
                print(json.dumps({"kernel": label, "ctas_per_sm": c, "ms": round(ms, 4),
                                  "input_GBps": round(10 * S / ms / 1e6, 1),
                                  "frac": round(14 * S / ms / 1e6 / peak, 4)}), flush=True)


if __name__ == "__main__":
    main()
