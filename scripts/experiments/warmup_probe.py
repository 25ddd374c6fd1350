#!/usr/bin/env python
"""scripts/experiments/warmup_probe.py — per-launch time of the shipped encode kernel from a cold process:
does the first ~100 ms after idle run slower (clock / memory power-state ramp)?  MEASUREMENT ONLY."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    import torch
    import seaweedfs_b200
    from seaweedfs_b200 import erasure_coding as ec
    L = seaweedfs_b200.lib()
    G = 1 << 30
    enc = ec.Encoder(10, 4, device=0)
    s = torch.cuda.current_stream().cuda_stream
    dat = torch.empty(30 * G, dtype=torch.uint8, device="cuda")
    par = [torch.empty(3 * G, dtype=torch.uint8, device="cuda") for _ in range(4)]
    pp = [p.data_ptr() for p in par]
    L.swec_synth_fill_device(0, dat.data_ptr(), 0, 30 * G, 7, s)
    torch.cuda.synchronize()
    for idle in (0.0, 2.0):
        time.sleep(idle)
        n = 120
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
        ev[0].record()
        for i in range(n):
            enc.encode_volume_device(dat.data_ptr(), 30 * G, pp, s)
            ev[i + 1].record()
        torch.cuda.synchronize()
        ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(n)]
        print(json.dumps({"idle_before_s": idle, "first_12": [round(x, 3) for x in ms[:12]],
                          "steps_13_24_mean": round(sum(ms[12:24]) / 12, 3), "steps_49_60_mean": round(sum(ms[48:60]) / 12, 3),
                          "steps_109_120_mean": round(sum(ms[108:120]) / 12, 3), "min": round(min(ms), 3)}), flush=True)


if __name__ == "__main__":
    main()
