#!/bin/bash
# rebuild the library on the GPU box with different load/store policies and sweep the leading shapes
OUT=gpurun_out/${1:-r1d}; mkdir -p $OUT
for flags in "-DSWEC_LD_POLICY=0" "-DSWEC_LD_POLICY=1" "-DSWEC_LD_POLICY=2" "-DSWEC_LD_POLICY=3"; do
  SWEC_EXTRA_NVCC_FLAGS="$flags" python seaweedfs_b200/build.py --force > /dev/null 2>> $OUT/build.err
  python scripts/tune_encode.py --best --steps 10 --tag="$flags" >> $OUT/policies.jsonl 2>> $OUT/tune.err
done
python seaweedfs_b200/build.py --force > /dev/null
cat $OUT/policies.jsonl
