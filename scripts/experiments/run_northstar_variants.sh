#!/bin/bash
# Build and run the north_star-sketch kernel next to the shipped one (GPU box; needs libswec.so built).
# usage: bash scripts/experiments/run_northstar_variants.sh [GiB] > profiles/rNN_northstar_variants.jsonl
set -e
cd "$(dirname "$0")/../.."
PEAK=$(python -c "import json;print(json.load(open('MEASURED_PEAKS.json'))['hbm_gbs'])" 2>/dev/null || echo 6650)
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Iinclude \
     -o /tmp/northstar_variants scripts/experiments/northstar_variants.cu \
     -Lseaweedfs_b200 -l:libswec.so -Xlinker -rpath -Xlinker "$PWD/seaweedfs_b200"
/tmp/northstar_variants "${1:-10}" "$PEAK"
