/*
 * oracle/rs_oracle.h — CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the reference's Reed–Solomon erasure-coding path,
 * used only as the checker by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg.  Nothing under seaweedfs_b200/ may include, link or call
 * this.  Every function cites the reference file:line it follows
 * (paths relative to /root/reference).
 *
 * Pinning status: PINNED — see oracle/README.md.  The oracle reproduces the
 * known-answer tests K1–K5 of the vendored reed-solomon-erasure crate, agrees
 * byte-for-byte with the reference's own compiled C kernel
 * (simd_c/reedsolomon.c → oracle/_ref/), and reproduces the SURVEY Appendix-A
 * digests for the reference fixture 1.dat.  (At the Go/klauspost boundary the
 * reference's own tests pin no parity bytes)
 */
#ifndef RS_ORACLE_H
#define RS_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_SHARDS 32 /* weed/storage/erasure_coding/ec_encoder.go:23 */

/* ---- GF(2^8), polynomial 0x11D (seaweed-volume/vendor/reed-solomon-erasure/build.rs:11-94) */
const uint8_t *orc_log_table(void);            /* [256]     build.rs:13-28  */
const uint8_t *orc_exp_table(void);            /* [510]     build.rs:30-42  */
const uint8_t *orc_mul_table(void);            /* [256*256] build.rs:58-68  */
void orc_mul_table_half(uint8_t c, uint8_t low[16], uint8_t high[16]); /* build.rs:70-94 */
uint8_t orc_mul(uint8_t a, uint8_t b);         /* src/galois_8.rs:67-69   */
uint8_t orc_div(uint8_t a, uint8_t b);         /* src/galois_8.rs:72-86   */
uint8_t orc_exp(uint8_t a, size_t n);          /* src/galois_8.rs:89-103  */
void orc_mul_slice(uint8_t c, const uint8_t *in, uint8_t *out, size_t n);     /* galois_8.rs:137-176 */
void orc_mul_slice_xor(uint8_t c, const uint8_t *in, uint8_t *out, size_t n); /* galois_8.rs:178-219 */

/* ---- matrices, row-major uint8 (src/matrix.rs) */
void orc_matrix_multiply(const uint8_t *a, int ar, int ac, const uint8_t *b, int bc, uint8_t *out); /* matrix.rs:119-139 */
int  orc_matrix_invert(const uint8_t *m, int n, uint8_t *out);  /* matrix.rs:195-261; 0 ok, -1 singular */
void orc_vandermonde(int rows, int cols, uint8_t *out);          /* matrix.rs:263-276 */
int  orc_build_matrix(int data_shards, int total_shards, uint8_t *out /* total*data */); /* core.rs:431-437 */

/* ---- codec (src/core.rs) ; shards[i] each n bytes */
int orc_encode(int k, int m, uint8_t *const *shards, size_t n);                  /* core.rs:600-635, 484-512 */
/* present[i]!=0 ⇒ shards[i] holds data; missing shards must point at n writable bytes. */
int orc_reconstruct(int k, int m, uint8_t *const *shards, const uint8_t *present,
                    size_t n, int data_only);                                   /* core.rs:736-926 */
int orc_verify(int k, int m, uint8_t *const *shards, size_t n);                  /* core.rs:640-672; 1 ok, 0 mismatch */
/* The k×k data-decode matrix for a given set of (first k) valid rows.  core.rs:700-734 */
int orc_decode_matrix(int k, int m, const uint8_t *present, uint8_t *out /* k*k */, int *valid /* k */);

/* ---- layout (weed/storage/erasure_coding) */
int64_t orc_expected_shard_size(int64_t dat_size, int k, int64_t large, int64_t small); /* weed/storage/disk_location_ec.go:428-448 */

typedef struct {
    int     block_index;
    int64_t inner_block_offset;
    int64_t size;
    int     is_large_block;
    int     large_block_rows_count;
} orc_interval;
/* ec_locate.go:16-53 ; returns number of intervals written (≤ cap), or -1 if cap too small */
int orc_locate_data(int64_t large, int64_t small, int64_t shard_dat_size, int64_t offset,
                    int64_t size, int k, orc_interval *out, int cap);
void orc_interval_to_shard(const orc_interval *iv, int64_t large, int64_t small, int k,
                           int *shard_id, int64_t *shard_offset);               /* ec_locate.go:87-98 */

/* ---- in-memory file walk: dat image → k+m shard images (ec_encoder.go:202-321).
 * shards[i] must hold orc_expected_shard_size() bytes. buffer_size as generateEcFiles. */
int orc_encode_dat_image(const uint8_t *dat, int64_t dat_size, int k, int m,
                         int64_t buffer_size, int64_t large, int64_t small,
                         uint8_t *const *shards);
/* shards .ec00-.ec(k-1) → dat image (ec_decoder.go:176-223) */
int orc_write_dat_image(uint8_t *dat, int64_t dat_size, int k, int64_t large, int64_t small,
                        const uint8_t *const *shards);

/* ---- file level (ec_encoder.go:110-128, 146-200, 323-377). 0 ok, negative errno-style */
int orc_generate_ec_files(const char *base, int64_t buffer_size, int64_t large, int64_t small,
                          int k, int m);
int orc_rebuild_ec_files(const char *base, int k, int m, uint32_t *rebuilt, int *nrebuilt);

/* ---- synthetic data: splitmix64 over the 8-byte word index. */
void orc_synth_fill(uint8_t *dst, uint64_t byte_offset, size_t n, uint64_t seed);

#ifdef __cplusplus
}
#endif
#endif
