#!/usr/bin/env python
"""scripts/experiments/power_policy_probe.py — MEASUREMENT ONLY: which kernel variant does the auto power policy pick
for the first launches of a process, a burst, and after a second of back-to-back encoding?  (SWEC_DEBUG_POWER=1 makes
the library print its decision for every Horner launch.)"""
import os
import sys
import time
os.environ["SWEC_DEBUG_POWER"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch                                                     # noqa: E402
import seaweedfs_b200                                            # noqa: E402
from seaweedfs_b200 import erasure_coding as ec                  # noqa: E402

enc = ec.Encoder(10, 4, device=0)
n = 1 << 30
d = [torch.zeros(n, dtype=torch.uint8, device="cuda") for _ in range(14)]
s = torch.cuda.current_stream().cuda_stream
print("== first encode", file=sys.stderr)
enc.encode_device([t.data_ptr() for t in d[:10]], [t.data_ptr() for t in d[10:]], n, s)
print("== reconstruct (AOT worst case) x3", file=sys.stderr)
for _ in range(3):
    enc.reconstruct_device([t.data_ptr() for t in d], [0, 0, 0, 0] + [1] * 10, n, False, s)
torch.cuda.synchronize()
print("== 400 encodes back to back (prints every 50th)", file=sys.stderr)
os.environ.pop("SWEC_DEBUG_POWER")
t0 = time.perf_counter()
for i in range(400):
    enc.encode_device([t.data_ptr() for t in d[:10]], [t.data_ptr() for t in d[10:]], n, s)
torch.cuda.synchronize()
print(f"== 400 x 10 GiB encodes took {time.perf_counter() - t0:.2f} s; one more:", file=sys.stderr)
enc.encode_device([t.data_ptr() for t in d[:10]], [t.data_ptr() for t in d[10:]], n, s)
time.sleep(2.0)
print("== after 2 s idle:", file=sys.stderr)
enc.encode_device([t.data_ptr() for t in d[:10]], [t.data_ptr() for t in d[10:]], n, s)
torch.cuda.synchronize()
