#!/usr/bin/env python
"""scripts/experiments/batch_gap_rows.py — MEASUREMENT ONLY: WHERE inside an encode launch does the 3-5 % go that a launch
loses when it follows another kernel (profiles/r02c_batch_probe.jsonl)?  The 30 GiB volume is encoded as three flat
launches, one per 10 x 1 GiB row, each bracketed by CUDA events, (a) back to back, (b) after a synth pass over the volume,
(c) after a 10 ms idle gap.  A start-up transient (clock ramp, cold L2 / TLB) shows in row 0 only; a steady-state effect
(DRAM state, power) in all three."""
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
GIB = 1 << 30


def main():
    import torch
    import seaweedfs_b200
    from seaweedfs_b200 import erasure_coding as ec
    L = seaweedfs_b200.lib()
    L.swec_set_option(b"power_mode", 1)                       # one variant throughout
    enc = ec.Encoder(10, 4, device=0)
    dat = torch.empty(30 * GIB, dtype=torch.uint8, device="cuda")
    par = [torch.empty(3 * GIB, dtype=torch.uint8, device="cuda") for _ in range(4)]
    s = torch.cuda.current_stream().cuda_stream
    L.swec_synth_fill_device(0, dat.data_ptr(), 0, 30 * GIB, 1, s)

    def rows():
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record()
        for r in range(3):
            d = [dat.data_ptr() + r * 10 * GIB + i * GIB for i in range(10)]
            p = [par[j].data_ptr() + r * GIB for j in range(4)]
            enc.encode_device(d, p, GIB, s)
            ev[r + 1].record()
        return ev

    def run(name, before):
        for _ in range(3):
            rows()
        torch.cuda.synchronize()
        per = [[], [], []]
        for v in range(30):
            before(v)
            ev = rows()
            ev[3].synchronize()
            for r in range(3):
                per[r].append(ev[r].elapsed_time(ev[r + 1]))
        print(json.dumps({"schedule": name, "row_ms_median": [round(statistics.median(x), 3) for x in per],
                          "row_ms_min": [round(min(x), 3) for x in per], "volume_ms_median": round(statistics.median(
                              [a + b + c for a, b, c in zip(*per)]), 3)}), flush=True)
        time.sleep(1.5)

    run("back_to_back", lambda v: None)
    run("after_synth", lambda v: L.swec_synth_fill_device(0, dat.data_ptr(), 0, 30 * GIB, 100 + v, s))
    run("after_10ms_idle", lambda v: (torch.cuda.synchronize(), time.sleep(0.010)))
    run("after_digest_of_parity", lambda v: [L.swec_digest_device(0, par[j].data_ptr(), 3 * GIB, __import__("ctypes").byref(__import__("ctypes").c_uint64(0)), s) for j in range(4)])


if __name__ == "__main__":
    main()
