"""The decode-kernel cache on the GPU, three tiers (include/swec.h swec_jit_stats): reconstruct matrices compiled with
the library (every single-shard loss + shards 0-3 lost) need neither NVRTC nor a warm-up; anything else is
specialised once and then loaded from the on-disk cubin cache by every later process — also by one that has no
NVRTC at all.  Reference behaviour mirrored: Encoder.Reconstruct keeps its decode matrices in a cache
(seaweed-volume/vendor/reed-solomon-erasure/src/core.rs:25,700-734; call site ec_encoder.go:360)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r"""
import ctypes as C, json, sys, time
import numpy as np
sys.path.insert(0, %(root)r)
import seaweedfs_b200
from oracle import pyoracle as po
from seaweedfs_b200 import erasure_coding as ec
import torch
L = seaweedfs_b200.lib()
enc = ec.Encoder(10, 4, device=0)
n = 8 << 20                                   # 80 MiB of input: above the inline-compile threshold
rng = np.random.default_rng(5)
data = [rng.integers(0, 256, n, dtype=np.uint8) for _ in range(10)]
full = data + po.encode(10, 4, data)
dev = [torch.from_numpy(x).cuda() for x in full]
out = {}
for name, lost in (("single_data", (6,)), ("single_parity", (12,)), ("worst", (0, 1, 2, 3)), ("two", (2, 11)), ("three", (1, 5, 13))):
    work = [t.clone() for t in dev]
    for i in lost:
        work[i].zero_()
    c0, h0 = C.c_uint64(0), C.c_uint64(0)
    L.swec_jit_stats(C.byref(c0), C.byref(h0), None, None)
    l0 = L.swec_kernel_launches()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    enc.reconstruct_device([t.data_ptr() for t in work], [0 if i in lost else 1 for i in range(14)], n, False,
                           torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    c1, h1 = C.c_uint64(0), C.c_uint64(0)
    L.swec_jit_stats(C.byref(c1), C.byref(h1), None, None)
    ok = all(torch.equal(work[i], dev[i]) for i in lost)
    out[name] = {"ok": ok, "compiles": c1.value - c0.value, "disk_hits": h1.value - h0.value,
                 "first_call_ms": round(dt * 1e3, 2), "launches": L.swec_kernel_launches() - l0}
print("RESULT " + json.dumps(out))
"""


def run_child(cache_dir, no_jit):
    env = dict(os.environ, SWEC_CACHE_DIR=str(cache_dir))
    env.pop("SWEC_NO_DISK_CACHE", None)
    if no_jit:
        env["SWEC_NO_JIT"] = "1"
    else:
        env.pop("SWEC_NO_JIT", None)
    r = subprocess.run([sys.executable, "-c", CHILD % {"root": ROOT}], env=env, stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    line = next(ln for ln in r.stdout.splitlines() if ln.startswith("RESULT "))
    return json.loads(line[7:])


def test_aot_patterns_need_no_nvrtc_and_disk_cache_serves_the_rest(cuda, swec, tmp_path):
    cache = tmp_path / "cubins"
    first = run_child(cache, no_jit=False)
    for name in ("single_data", "single_parity", "worst"):          # compiled with the library
        assert first[name] == dict(first[name], ok=True, compiles=0, disk_hits=0, launches=1), (name, first[name])
    for name in ("two", "three"):                                     # first sight: NVRTC, then into the cache
        assert first[name]["ok"] and first[name]["compiles"] == 1 and first[name]["disk_hits"] == 0, (name, first[name])
    assert len(list(cache.glob("*.cubin"))) == 2
    # a second process WITHOUT NVRTC: AOT patterns unchanged, the other two come from the disk cache in milliseconds
    second = run_child(cache, no_jit=True)
    for name in ("single_data", "single_parity", "worst"):
        assert second[name]["ok"] and second[name]["compiles"] == 0 and second[name]["launches"] == 1
    for name in ("two", "three"):
        assert second[name]["ok"] and second[name]["compiles"] == 0 and second[name]["disk_hits"] == 1, (name, second[name])
        assert second[name]["launches"] == 1                          # the specialised kernel, not the table kernel
        assert second[name]["first_call_ms"] < first[name]["first_call_ms"], (first[name], second[name])
    print(json.dumps({"first_process": first, "second_process_no_nvrtc": second}))
    # without NVRTC and without the cache the engine still answers — from the shared-memory table kernel
    third = run_child(tmp_path / "empty", no_jit=True)
    assert third["two"]["ok"] and third["two"]["compiles"] == 0 and third["two"]["disk_hits"] == 0
