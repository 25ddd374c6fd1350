#!/usr/bin/env python
"""scripts/pcie_socket_probe.py — what can the host feed to 1, 2, 4 GPUs of ONE socket (and to all 8) at once?

bench.py's e2e leg at N=4/8 reaches 0.64-0.66 of the per-GPU H2D rate a single GPU gets alone (SCALE_r01.json):
four GPUs behind one socket saturate at ~142 GB/s in + ~57 GB/s out.  This probe names the bound: plain DMA, no
kernels, no library code besides the pinned allocator — per GPU a NUMA-local pinned buffer (swec_alloc_pinned_for_device,
huge pages unless SWEC_NO_THP=1), H2D-only, D2H-only and both directions at once, for GPU sets {1 of socket 0},
{2 of socket 0}, {4 of socket 0}, {one per socket}, {2+2}, {all}.  Device-timed (CUDA events per GPU), aggregate =
sum of per-GPU bytes / max per-GPU time.  One JSON line per (set, direction)."""
import ctypes as C
import json
import os
import sys
import threading

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GIB = 1 << 30


def main():
    import numpy as np
    import torch
    import seaweedfs_b200
    L = seaweedfs_b200.lib()
    ng = torch.cuda.device_count()
    node = {}
    for g in range(ng):
        bus = torch.cuda.get_device_properties(g).pci_bus_id if hasattr(torch.cuda.get_device_properties(g), "pci_bus_id") else None
        n = -1
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(g)
            bus = pynvml.nvmlDeviceGetPciInfo(h).busId
            bus = bus.decode() if isinstance(bus, bytes) else bus
            n = int(open(f"/sys/bus/pci/devices/{bus[-12:].lower()}/numa_node").read())
        except Exception:
            pass
        node[g] = n
    size = int(float(os.environ.get("PROBE_GIB", "4")) * GIB)
    host, dev_in, dev_out, raws = {}, {}, {}, {}
    for g in range(ng):
        raw = L.swec_alloc_pinned_for_device(g, 2 * size)
        assert raw
        raws[g] = raw
        arr = np.ctypeslib.as_array(C.cast(raw, C.POINTER(C.c_uint8)), shape=(2 * size,))
        arr[::4096] = 1                                            # touch
        host[g] = torch.from_numpy(arr)
        with torch.cuda.device(g):
            dev_in[g] = torch.empty(size, dtype=torch.uint8, device=f"cuda:{g}")
            dev_out[g] = torch.ones(size, dtype=torch.uint8, device=f"cuda:{g}")
    s0 = [g for g in range(ng) if node[g] == node[0]]
    s1 = [g for g in range(ng) if node[g] != node[0]]
    sets = [("1 GPU", s0[:1])]
    if len(s0) >= 2:
        sets.append(("2 GPUs, one socket", s0[:2]))
    if len(s0) >= 4:
        sets.append(("4 GPUs, one socket", s0[:4]))
    if s1:
        sets.append(("2 GPUs, one per socket", [s0[0], s1[0]]))
    if len(s0) >= 2 and len(s1) >= 2:
        sets.append(("4 GPUs, two per socket", s0[:2] + s1[:2]))
    if len(s0) + len(s1) >= 8:
        sets.append(("8 GPUs", s0[:4] + s1[:4]))
    reps = 3

    def run(gpus, h2d, d2h):
        times = {}
        bar = threading.Barrier(len(gpus))

        def one(g):
            with torch.cuda.device(g):
                si, so = torch.cuda.Stream(), torch.cuda.Stream()
                for timed in (False, True):
                    torch.cuda.synchronize()
                    bar.wait()
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record()
                    for _ in range(reps if timed else 1):
                        if h2d:
                            with torch.cuda.stream(si):
                                dev_in[g].copy_(host[g][:size], non_blocking=True)
                        if d2h:
                            with torch.cuda.stream(so):
                                host[g][size:].copy_(dev_out[g], non_blocking=True)
                    torch.cuda.current_stream().wait_stream(si)
                    torch.cuda.current_stream().wait_stream(so)
                    b.record()
                    torch.cuda.synchronize()
                    times[g] = a.elapsed_time(b) / 1e3
        th = [threading.Thread(target=one, args=(g,)) for g in gpus]
        [t.start() for t in th]
        [t.join() for t in th]
        tmax = max(times.values())
        return tmax, times

    for name, gpus in sets:
        for dname, h2d, d2h in (("h2d", True, False), ("d2h", False, True), ("both", True, True)):
            tmax, times = run(gpus, h2d, d2h)
            per = reps * size / 1e9
            row = {"set": name, "gpus": gpus, "numa_nodes": [node[g] for g in gpus], "direction": dname,
                   "h2d_GBps": round(len(gpus) * per / tmax, 1) if h2d else 0,
                   "d2h_GBps": round(len(gpus) * per / tmax, 1) if d2h else 0,
                   "per_gpu_GBps_each_direction": [round(per / times[g], 1) for g in gpus],
                   "huge_pages": not os.environ.get("SWEC_NO_THP"), "buffer_GiB_per_gpu_per_direction": size / GIB}
            print(json.dumps(row), flush=True)
    for g in range(ng):
        L.swec_free_pinned(raws[g])


if __name__ == "__main__":
    main()
