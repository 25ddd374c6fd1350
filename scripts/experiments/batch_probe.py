#!/usr/bin/env python
"""scripts/experiments/batch_probe.py — why does an encode launch inside the 256-volume batch take ~7.7 ms when the
same launch takes 6.9-7.2 ms back to back?  MEASUREMENT ONLY.  Per schedule, 40 volumes; per volume the encode launch
is timed with CUDA events and the SM clock / power are read through NVML right after it:
  encode_only          encode, encode, …                                  (the sustained leg)
  synth_encode_digest  synth(30 GiB) → encode → 4 digests                  (the batch leg)
  synth_gap_encode     synth → sync → 5 ms host sleep → encode             (does the preceding kernel matter?)
  digest_encode        4 digests → encode
  memset_encode        cudaMemset of the volume (DMA engine fill, no SM work) → encode
Prints one JSON line per schedule: median/mean encode ms, fraction of the HBM peak, clocks, power."""
import ctypes as C
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
    import pynvml
    import seaweedfs_b200
    from seaweedfs_b200 import erasure_coding as ec
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    L = seaweedfs_b200.lib()
    enc = ec.Encoder(10, 4, device=0)
    dat_size = 30 * GIB
    shard = ec.expected_shard_size(dat_size)
    dat = torch.empty(dat_size, dtype=torch.uint8, device="cuda")
    par = [torch.empty(shard, dtype=torch.uint8, device="cuda") for _ in range(4)]
    pp = [p.data_ptr() for p in par]
    s = torch.cuda.current_stream().cuda_stream
    L.swec_synth_fill_device(0, dat.data_ptr(), 0, dat_size, 1, s)
    peak = 6501.2
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass

    def synth(v):
        assert L.swec_synth_fill_device(0, dat.data_ptr(), 0, dat_size, 100 + v, s) == 0

    def digests():
        one = C.c_uint64(0)
        for p in range(4):
            assert L.swec_digest_device(0, pp[p], shard, C.byref(one), s) == 0

    def run(name, before, after):
        for _ in range(5):
            enc.encode_volume_device(dat.data_ptr(), dat_size, pp, s)
        torch.cuda.synchronize()
        ms, clk, pw = [], [], []
        t0 = time.perf_counter()
        for v in range(40):
            before(v)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            enc.encode_volume_device(dat.data_ptr(), dat_size, pp, s)
            b.record()
            after(v)
            b.synchronize()
            clk.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
            pw.append(pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0)
            ms.append(a.elapsed_time(b))
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        med = statistics.median(ms)
        print(json.dumps({"schedule": name, "encode_ms_median": round(med, 3), "encode_ms_mean": round(sum(ms) / len(ms), 3),
                          "encode_ms_min": round(min(ms), 3), "encode_ms_last10": [round(x, 2) for x in ms[-10:]],
                          "roofline_frac_median": round(1.4 * dat_size / (med / 1e3) / 1e9 / peak, 4),
                          "sm_mhz_median": statistics.median(clk), "power_w_median": round(statistics.median(pw), 1),
                          "wall_s": round(wall, 2), "encode_duty": round(sum(ms) / 1e3 / wall, 3)}), flush=True)
        time.sleep(2.0)

    nop = lambda v: None                                                    # noqa: E731
    run("encode_only", nop, nop)
    run("synth_encode_digest", synth, lambda v: digests())
    run("synth_gap_encode", lambda v: (synth(v), torch.cuda.synchronize(), time.sleep(0.005)), nop)
    run("digest_encode", lambda v: digests(), nop)
    run("memset_encode", lambda v: torch.cuda.synchronize() or C.c_int(0) and None or dat.zero_(), nop)
    for mode in (1, 2):
        L.swec_set_option(b"power_mode", mode)
        run(f"synth_encode_digest_power_mode_{mode}", synth, lambda v: digests())
    L.swec_set_option(b"power_mode", 0)


if __name__ == "__main__":
    main()
