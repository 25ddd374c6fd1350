#!/usr/bin/env python
"""scripts/experiments/stride_probe.py — does the distance between the ten input streams matter to HBM?
Same kernel body (rs10x4_encode, flat), same bytes (10 x 3 GiB in, 4 x 3 GiB out); only the stream base
addresses move: exactly 1 GiB / 3 GiB apart (what the .dat layout gives), or de-aligned by odd multiples of
a few KiB.  Also the blocked whole-volume launch for reference.  MEASUREMENT ONLY."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    import torch
    import seaweedfs_b200
    from seaweedfs_b200 import erasure_coding as ec
    L = seaweedfs_b200.lib()
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    G = 1 << 30
    enc = ec.Encoder(10, 4, device=0)
    s = torch.cuda.current_stream().cuda_stream
    pool = torch.empty(34 * G, dtype=torch.uint8, device="cuda")
    L.swec_synth_fill_device(0, pool.data_ptr(), 0, 34 * G, 7, s)
    par = [torch.empty(3 * G + (1 << 20), dtype=torch.uint8, device="cuda") for _ in range(4)]
    base = pool.data_ptr()

    def timed(fn, steps=8):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / steps

    def report(label, ms, nbytes_in):
        print(json.dumps({"case": label, "ms": round(ms, 4), "input_GBps": round(nbytes_in / ms / 1e6, 1),
                          "frac": round(1.4 * nbytes_in / ms / 1e6 / peak, 4)}), flush=True)

    pp = [p.data_ptr() for p in par]
    for rep in range(2):
        ms = timed(lambda: enc.encode_volume_device(base, 30 * G, pp, s))
        report("blocked volume, 3 rows of 10 x 1 GiB (shipped)", ms, 30 * G)
        n = 3 * G
        for label, stride, ostride in (("flat, streams 3 GiB apart", 3 * G, 0),
                                       ("flat, streams 3 GiB + 12 KiB apart", 3 * G + 12288, 0),
                                       ("flat, streams 3 GiB + 68 KiB apart", 3 * G + 69632, 0),
                                       ("flat, streams 3 GiB + 1 MiB + 4 KiB apart", 3 * G + (1 << 20) + 4096, 0),
                                       ("flat, streams 3 GiB + 68 KiB apart, outputs shifted i x 36 KiB", 3 * G + 69632, 36864)):
            d = [base + i * stride for i in range(10)]
            o = [pp[j] + j * ostride for j in range(4)]
            ms = timed(lambda: enc.encode_device(d, o, n, s))
            report(label, ms, 10 * n)
        # one row only (10 GiB), blocked-equivalent addresses but flat: streams exactly 1 GiB apart vs padded
        n1 = G
        for label, stride in (("flat 1 GiB streams exactly 1 GiB apart", G), ("flat 1 GiB streams 1 GiB + 68 KiB apart", G + 69632)):
            d = [base + i * stride for i in range(10)]
            ms = timed(lambda: enc.encode_device(d, pp, n1, s), steps=16)
            report(label, ms, 10 * n1)


if __name__ == "__main__":
    main()
