#!/usr/bin/env python
"""scripts/bench_decode.py DIR GIB — WriteDatFile (ec.decode's copy of .ec00-.ec09 back into a .dat, ec_decoder.go:176-223)
through swec_write_dat_file: pieces with explicit offsets copied in parallel (copy_file_range, else pread/pwrite).
SWEC_IO_THREADS=1 approximates the reference's single sequential io.CopyN stream.  No GPU involved."""
import os, sys, time, numpy as np, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import seaweedfs_b200
from seaweedfs_b200 import erasure_coding as ec
from oracle import pyoracle as po
d=sys.argv[1]; gib=float(sys.argv[2])
base=os.path.join(d,'wdf'); size=int(gib*(1<<30))+777
rng=np.random.default_rng(1)
blk=rng.integers(0,256,64<<20,dtype=np.uint8)
k=10; small=1<<20
shard=ec.expected_shard_size(size)
# write data shards directly: shard i = concatenation of its blocks (content irrelevant for timing; use pattern)
for i in range(k):
    with open(base+'.ec%02d'%i,'wb') as f:
        left=shard
        while left>0:
            n=min(left,len(blk)); f.write(blk[:n].tobytes()); left-=n
for rep in range(2):
    t0=time.perf_counter(); ec.WriteDatFile(base+'_out', size, [base+'.ec%02d'%i for i in range(k)]); dt=time.perf_counter()-t0
    print(d, 'WriteDatFile %.2f GiB: %.3f s = %.2f GB/s'%(gib, dt, size/dt/1e9))
# spot check bytes: first small row
out=np.memmap(base+'_out.dat',dtype=np.uint8,mode='r')
s0=np.fromfile(base+'.ec00',dtype=np.uint8,count=small); s1=np.fromfile(base+'.ec01',dtype=np.uint8,count=small)
assert (out[:small]==s0).all() and (out[small:2*small]==s1).all() and len(out)==size
for f in os.listdir(d):
    if f.startswith('wdf'): os.remove(os.path.join(d,f))
