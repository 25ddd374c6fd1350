#!/usr/bin/env python
"""scripts/bench_host_api.py — sweep of the Encoder-level seam from host memory (bench.py's `host_api` leg) over the
data paths the engine has for it: DMA ring vs zero-copy kernel, number of pipelined pieces per call, smallest piece.
One JSON line per configuration; the 1-thread CPU arithmetic on the same buffers rides in every line."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import seaweedfs_b200
    from seaweedfs_b200 import erasure_coding as ec
    import bench
    L = seaweedfs_b200.lib()
    enc = ec.Encoder(10, 4, device=0)
    sizes = (64 * 1024, 256 * 1024, 1 << 20, 4 << 20, 16 << 20)
    configs = [(zc, pieces, minc) for zc in (0, 1) for pieces in (1, 2, 4) for minc in (64 << 10, 256 << 10)]
    if len(sys.argv) > 1 and sys.argv[1] == "--quick":
        configs = [(0, 4, 128 << 10), (1, 1, 128 << 10), (1, 2, 128 << 10), (1, 4, 128 << 10)]
    for zc, pieces, minc in configs:
        assert L.swec_set_option(b"host_zero_copy", zc) == 0
        assert L.swec_set_option(b"host_pieces", pieces) == 0
        assert L.swec_set_option(b"host_min_chunk", minc) == 0
        leg = bench.host_api_leg(L, enc, 0, sizes=sizes, min_s=0.25)
        row = {"zero_copy": zc, "pieces": pieces, "min_chunk": minc}
        for n, r in leg["sizes"].items():
            row[n] = {k.replace("_encode_GBps", "").replace("_GBps", ""): v for k, v in r.items() if k.endswith("GBps")}
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
