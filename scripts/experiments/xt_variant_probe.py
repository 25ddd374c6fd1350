#!/usr/bin/env python
"""scripts/experiments/xt_variant_probe.py — instruction-mix variants of the multiply-by-2 step
(device_common.cuh, SWEC_XT_VARIANT) under BURST and SUSTAINED load.  The encode matrix is specialised at run
time for each variant ("use_aot" 0), every variant's parity is compared with the AOT kernel's digests, and each
is timed over 150 back-to-back 30 GiB encodes: mean of launches 11-30 (boost clocks) and 91-150 (the GPU has
settled against its power cap).  MEASUREMENT ONLY."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, ROOT)


def main():
    import torch
    import seaweedfs_b200
    from bench import ClockSampler
    from seaweedfs_b200 import erasure_coding as ec
    L = seaweedfs_b200.lib()
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    G = 1 << 30
    s = torch.cuda.current_stream().cuda_stream
    dat = torch.empty(30 * G, dtype=torch.uint8, device="cuda")
    par = [torch.empty(3 * G, dtype=torch.uint8, device="cuda") for _ in range(4)]
    pp = [p.data_ptr() for p in par]
    L.swec_synth_fill_device(0, dat.data_ptr(), 0, 30 * G, 7, s)

    def digests():
        out = []
        for p in pp:
            d = C.c_uint64(0)
            assert L.swec_digest_device(0, p, 3 * G, C.byref(d), s) == 0
            out.append(d.value)
        return out

    ref = None
    for rep in range(1):
        for label, aot, variant in (("aot v0", 1, 0), ("jit v2", 0, 2), ("jit v3", 0, 3)):
            assert L.swec_set_option(b"use_aot", aot) == 0 and L.swec_set_option(b"xt_variant", variant) == 0
            enc = ec.Encoder(10, 4, device=0)
            for p in par:
                p.zero_()
            enc.encode_volume_device(dat.data_ptr(), 30 * G, pp, s)       # compiles on first use
            got = digests()
            ref = ref or got
            n = 300
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
            with ClockSampler(0, None) as clk:
                ev[0].record()
                for i in range(n):
                    enc.encode_volume_device(dat.data_ptr(), 30 * G, pp, s)
                    ev[i + 1].record()
                torch.cuda.synchronize()
            ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(n)]
            burst, sustained = sum(ms[10:30]) / 20, sum(ms[200:300]) / 100
            c = clk.summary()
            print(json.dumps({"kernel": label, "bit_exact_vs_aot": got == ref, "burst_ms": round(burst, 3),
                              "burst_frac": round(1.4 * 30 * G / burst / 1e6 / peak, 4), "sustained_ms": round(sustained, 3),
                              "sustained_frac": round(1.4 * 30 * G / sustained / 1e6 / peak, 4),
                              "power_w_max": c["power_w_max"], "sm_mhz_min": c["sm_min_mhz"], "reasons": c["reasons"],
                              "ms_every_5th_launch": [round(x, 2) for x in ms[::5]],
                              "sm_mhz_timeline": [smp[0] for smp in clk.samples[::max(1, len(clk.samples) // 30)]],
                              "power_w_timeline": [round(smp[2]) for smp in clk.samples[::max(1, len(clk.samples) // 30)]]}), flush=True)
            enc.close()
            torch.cuda.synchronize()
            import time
            time.sleep(3)                                                  # let the package cool between variants


if __name__ == "__main__":
    main()
