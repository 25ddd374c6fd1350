#!/usr/bin/env python
"""tests/golden/make_golden.py — regenerates tests/golden/kat.json.  Run in the build container
(needs /root/reference); the JSON it writes is committed and is all the GPU box ever reads.

Sources (paths under /root/reference; "rse/" = seaweed-volume/vendor/reed-solomon-erasure/):
  K1 log table          rse/src/galois_8.rs:339-364   (parsed from the Rust test source)
  K2 scalar mul/exp     rse/src/galois_8.rs:483-485,549-551
  K3 slice vectors      rse/src/galois_8.rs:487-547   (parsed)
  K4 matrix inverse     rse/src/matrix.rs:382-413
  K5 RS(5,5) encode     rse/src/tests/mod.rs:851-893
  K7 LocateData         weed/storage/erasure_coding/ec_test.go:199-274
  K8 fixture digests    weed/storage/erasure_coding/1.dat encoded by the REFERENCE'S OWN compiled C
                        kernel (rse/simd_c/reedsolomon.c → oracle/_ref/libref_rs_*.so) driven like
                        code_some_slices (rse/src/core.rs:484-512) over the shard layout of
                        encodeDatFile (ec_encoder.go:280-321), production and test block sizes
  K9 patterns           weed/storage/store_ec_recovery_test.go:208-213, ec_encoder.rs:666-674 (inputs;
                        parity bytes computed through the reference kernel as for K8)
"""
import hashlib
import json
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference"
RSE = REF + "/seaweed-volume/vendor/reed-solomon-erasure"

from oracle import pyoracle as po  # noqa: E402
from oracle import rs_numpy as rn  # noqa: E402


def ints(text):
    return [int(x, 0) for x in re.findall(r"0x[0-9a-fA-F]+|\d+", text)]


def main():
    po.build(quiet=True)
    assert po.ref_available(), "oracle/_ref missing"
    g8 = open(RSE + "/src/galois_8.rs").read()
    out = {}
    m = re.search(r"BACKBLAZE_LOG_TABLE: \[u8; 256\] = \[(.*?)\];", g8, re.S)
    body = re.sub(r"//.*", "", m.group(1))
    out["K1_log_table"] = ints(body)
    assert len(out["K1_log_table"]) == 256
    out["K2_mul"] = [[3, 4, 12], [7, 7, 21], [23, 45, 41]]
    out["K2_exp"] = [[2, 2, 4], [5, 20, 235], [13, 7, 43]]
    t = g8[g8.index("fn test_galois()"):g8.index("fn test_slice_add()")]
    arrays = [ints(a) for a in re.findall(r"= \[(.*?)\];", t, re.S)]
    assert len(arrays) == 5 and all(len(a) == 34 for a in arrays)
    out["K3"] = {"input": arrays[0], "mul_25": arrays[1], "then_xor_52": arrays[2],
                 "mul_177": arrays[3], "then_xor_117": arrays[4]}
    out["K4"] = {"m": [[56, 23, 98], [3, 100, 200], [45, 201, 123]],
                 "inv": [[175, 133, 33], [130, 13, 245], [112, 35, 126]],
                 "m5": [[1, 0, 0, 0, 0], [0, 1, 0, 0, 0], [0, 0, 0, 1, 0], [0, 0, 0, 0, 1], [7, 7, 6, 6, 1]],
                 "inv5": [[1, 0, 0, 0, 0], [0, 1, 0, 0, 0], [123, 123, 1, 122, 122], [0, 0, 1, 0, 0], [0, 0, 0, 1, 0]]}
    out["K5"] = {"data": [[0, 1], [4, 5], [2, 3], [6, 7], [8, 9]],
                 "parity": [[12, 13], [10, 11], [14, 15], [90, 91], [94, 95]]}
    out["K7"] = [
        {"args": [1 << 30, 1 << 20, 3221225472 - 1, 21479557912, 4194339],
         "intervals": [[4, 527128, 521448, False, 2], [5, 0, 1048576, False, 2], [6, 0, 1048576, False, 2],
                       [7, 0, 1048576, False, 2], [8, 0, 527163, False, 2]]},
        {"args": [1 << 30, 1 << 20, 3221225472 - 1, 30782909808, 112568],
         "intervals": [[8876, 912752, 112568, False, 2]]},
        {"args": [10000, 100, 10001, 100000, 1], "intervals": [[0, 0, 1, False, 1]]},
        {"args": [1 << 30, 1 << 20, 3 << 30, 20 << 30, 1024], "intervals": [[20, 0, 1024, True, 3]]},
    ]

    # K8 / K9 through the reference's own compiled kernel
    gen = rn.build_matrix(10, 14)

    def ref_parity(data):
        n = len(data[0])
        outs = [np.zeros(n, dtype=np.uint8) for _ in range(4)]
        po.cpu_apply(0, gen[10:], [np.ascontiguousarray(d) for d in data], outs, threads=1)
        return outs

    dat = np.fromfile(REF + "/weed/storage/erasure_coding/1.dat", dtype=np.uint8)
    out["K8"] = {"dat_sha256": hashlib.sha256(dat.tobytes()).hexdigest(), "dat_size": int(dat.shape[0])}
    for label, (large, small) in {"production": (1 << 30, 1 << 20), "test": (10000, 100)}.items():
        shards = rn.encode_dat_image(dat, 10, 4, large, small)[:10]  # layout only
        shards += ref_parity(shards)                                  # arithmetic by the reference kernel
        out["K8"][label] = {"large": large, "small": small, "shard_size": int(shards[0].shape[0]),
                            "sha256": [hashlib.sha256(s.tobytes()).hexdigest() for s in shards]}
    n = 64
    pat_a = [np.full(n, (7 * i) & 255, dtype=np.uint8) for i in range(10)]
    pat_b = [((np.arange(n) + i) & 255).astype(np.uint8) for i in range(10)]
    out["K9"] = {"seven_i": [int(p[0]) for p in ref_parity(pat_a)],
                 "i_plus_j": [[int(p[j]) for p in ref_parity(pat_b)] for j in range(4)]}
    out["generator_rs10_4_parity_rows"] = gen[10:].tolist()
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "kat.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote kat.json")


if __name__ == "__main__":
    main()
