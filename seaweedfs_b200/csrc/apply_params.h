// seaweedfs_b200/csrc/apply_params.h — kernel launch parameters shared by host code, the AOT
// kernels and the NVRTC-specialised kernels (the build inlines this file into the JIT prelude, so
// it must stay free of #includes).
#pragma once

typedef unsigned char u8;
typedef unsigned int u32;
typedef unsigned long long u64;

#define SWEC_MAX_INPUTS 32   // MaxShardCount, weed/storage/erasure_coding/ec_encoder.go:23
#define SWEC_MAX_OUTPUTS 8   // outputs per launch; more rows → more launches

// Flat layout: stream i is in[i][0 .. 16*nvec).  Blocked layout (the .dat striping of
// encodeDatFile, ec_encoder.go:280-321): the image is rows of K blocks; in[i] points at block i of
// row 0 and consecutive rows are K*block bytes apart, while outputs are contiguous.  With
// v = row*block_vecs + xv the input offset is 16*v + row*row_extra, row_extra = (K-1)*block.
struct SwecApplyParams {
    const u8* in[SWEC_MAX_INPUTS];
    u8* out[SWEC_MAX_OUTPUTS];
    u64 nvec;        // 16-byte vectors per stream
    u64 block_vecs;  // vectors per block (blocked layout only)
    u64 row_extra;   // bytes
    int block_shift; // log2(block_vecs) or -1
};

