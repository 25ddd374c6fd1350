/*
 * oracle/rs_oracle.c — CPU ORACLE (test infrastructure, NOT product code).
 * See rs_oracle.h for the contract and pinning status.  Scalar, single-threaded,
 * written for clarity: every routine restates one reference routine and names it.
 * Paths are relative to /root/reference; "rse/" abbreviates
 * seaweed-volume/vendor/reed-solomon-erasure/.
 */
#define _GNU_SOURCE
#include "rs_oracle.h"

#include <errno.h>
#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

/* ------------------------------------------------------------------ field */

enum { FIELD = 256, POLY = 29 /* rse/build.rs:11 — 0x11D without x^8 */ };

static uint8_t g_log[FIELD];
static uint8_t g_exp[2 * FIELD - 2];
static uint8_t g_mul[FIELD * FIELD];
static int g_ready;

/* rse/build.rs:13-28 (log), :30-42 (exp), :58-68 (mul) */
static void tables_init(void)
{
    if (g_ready) return;
    unsigned b = 1;
    for (unsigned lg = 0; lg < FIELD - 1; lg++) {
        g_log[b] = (uint8_t)lg;
        b <<= 1;
        if (b >= FIELD) b = (b - FIELD) ^ POLY;
    }
    for (unsigned i = 1; i < FIELD; i++) {
        unsigned lg = g_log[i];
        g_exp[lg] = (uint8_t)i;
        g_exp[lg + FIELD - 1] = (uint8_t)i;
    }
    for (unsigned a = 0; a < FIELD; a++)
        for (unsigned c = 0; c < FIELD; c++)
            g_mul[a * FIELD + c] = (a == 0 || c == 0) ? 0 : g_exp[(unsigned)g_log[a] + g_log[c]];
    g_ready = 1;
}

const uint8_t *orc_log_table(void) { tables_init(); return g_log; }
const uint8_t *orc_exp_table(void) { tables_init(); return g_exp; }
const uint8_t *orc_mul_table(void) { tables_init(); return g_mul; }

/* rse/build.rs:70-94: low[b] = c*b for b<16 ; high[b>>4] = c*b for b = multiples of 16 */
void orc_mul_table_half(uint8_t c, uint8_t low[16], uint8_t high[16])
{
    tables_init();
    for (unsigned n = 0; n < 16; n++) {
        low[n] = g_mul[c * FIELD + n];
        high[n] = g_mul[c * FIELD + (n << 4)];
    }
}

uint8_t orc_mul(uint8_t a, uint8_t b) { tables_init(); return g_mul[a * FIELD + b]; }

/* rse/src/galois_8.rs:72-86 */
uint8_t orc_div(uint8_t a, uint8_t b)
{
    tables_init();
    if (a == 0) return 0;
    if (b == 0) abort(); /* reference panics */
    int d = (int)g_log[a] - (int)g_log[b];
    if (d < 0) d += 255;
    return g_exp[d];
}

/* rse/src/galois_8.rs:89-103 */
uint8_t orc_exp(uint8_t a, size_t n)
{
    tables_init();
    if (n == 0) return 1;
    if (a == 0) return 0;
    size_t lg = (size_t)g_log[a] * n;
    lg %= 255; /* the reference subtracts 255 until < 255 */
    return g_exp[lg];
}

/* rse/src/galois_8.rs:137-176 (pure path): out = c ⊗ in */
void orc_mul_slice(uint8_t c, const uint8_t *in, uint8_t *out, size_t n)
{
    tables_init();
    const uint8_t *row = &g_mul[c * FIELD];
    for (size_t i = 0; i < n; i++) out[i] = row[in[i]];
}

/* rse/src/galois_8.rs:178-219: out ^= c ⊗ in */
void orc_mul_slice_xor(uint8_t c, const uint8_t *in, uint8_t *out, size_t n)
{
    tables_init();
    const uint8_t *row = &g_mul[c * FIELD];
    for (size_t i = 0; i < n; i++) out[i] ^= row[in[i]];
}

/* ----------------------------------------------------------------- matrix */

/* rse/src/matrix.rs:119-139 */
void orc_matrix_multiply(const uint8_t *a, int ar, int ac, const uint8_t *b, int bc, uint8_t *out)
{
    for (int r = 0; r < ar; r++)
        for (int c = 0; c < bc; c++) {
            uint8_t v = 0;
            for (int i = 0; i < ac; i++) v ^= orc_mul(a[r * ac + i], b[i * bc + c]);
            out[r * bc + c] = v;
        }
}

/* rse/src/matrix.rs:195-261: augment with I, Gauss–Jordan, take the right half */
int orc_matrix_invert(const uint8_t *m, int n, uint8_t *out)
{
    int w = 2 * n;
    uint8_t *t = (uint8_t *)calloc((size_t)n * w, 1);
    if (!t) return -2;
    for (int r = 0; r < n; r++) {
        memcpy(&t[r * w], &m[r * n], (size_t)n);
        t[r * w + n + r] = 1;
    }
    for (int r = 0; r < n; r++) {
        if (t[r * w + r] == 0) {
            for (int rb = r + 1; rb < n; rb++)
                if (t[rb * w + r] != 0) {
                    for (int c = 0; c < w; c++) {
                        uint8_t x = t[r * w + c];
                        t[r * w + c] = t[rb * w + c];
                        t[rb * w + c] = x;
                    }
                    break;
                }
        }
        if (t[r * w + r] == 0) { free(t); return -1; }
        if (t[r * w + r] != 1) {
            uint8_t s = orc_div(1, t[r * w + r]);
            for (int c = 0; c < w; c++) t[r * w + c] = orc_mul(s, t[r * w + c]);
        }
        for (int rb = r + 1; rb < n; rb++) {
            uint8_t s = t[rb * w + r];
            if (s)
                for (int c = 0; c < w; c++) t[rb * w + c] ^= orc_mul(s, t[r * w + c]);
        }
    }
    for (int d = 0; d < n; d++)
        for (int ra = 0; ra < d; ra++) {
            uint8_t s = t[ra * w + d];
            if (s)
                for (int c = 0; c < w; c++) t[ra * w + c] ^= orc_mul(s, t[d * w + c]);
        }
    for (int r = 0; r < n; r++) memcpy(&out[r * n], &t[r * w + n], (size_t)n);
    free(t);
    return 0;
}

/* rse/src/matrix.rs:263-276: entry (r,c) = r^c, 0^0 = 1 */
void orc_vandermonde(int rows, int cols, uint8_t *out)
{
    for (int r = 0; r < rows; r++)
        for (int c = 0; c < cols; c++) out[r * cols + c] = orc_exp((uint8_t)r, (size_t)c);
}

/* rse/src/core.rs:431-437: vandermonde(total,data) × inverse(top data×data) */
int orc_build_matrix(int k, int total, uint8_t *out)
{
    if (k <= 0 || total <= k || total > 256) return -1;
    uint8_t *v = (uint8_t *)malloc((size_t)total * k);
    uint8_t *top_inv = (uint8_t *)malloc((size_t)k * k);
    if (!v || !top_inv) { free(v); free(top_inv); return -2; }
    orc_vandermonde(total, k, v);
    int rc = orc_matrix_invert(v, k, top_inv); /* first k rows of v are the top square */
    if (rc == 0) orc_matrix_multiply(v, total, k, top_inv, k, out);
    free(v);
    free(top_inv);
    return rc;
}

/* ------------------------------------------------------------------ codec */

/* rse/src/core.rs:484-512: for each input, for each output row: first input mul, rest mul_xor */
static void code_some_slices(const uint8_t *const *rows, int nrows, int k,
                             const uint8_t *const *inputs, uint8_t *const *outputs, size_t n)
{
    for (int i = 0; i < k; i++)
        for (int r = 0; r < nrows; r++) {
            if (i == 0) orc_mul_slice(rows[r][i], inputs[i], outputs[r], n);
            else orc_mul_slice_xor(rows[r][i], inputs[i], outputs[r], n);
        }
}

/* rse/src/core.rs:600-635 */
int orc_encode(int k, int m, uint8_t *const *shards, size_t n)
{
    if (k <= 0 || m <= 0 || k + m > 256) return -1;
    uint8_t *mat = (uint8_t *)malloc((size_t)(k + m) * k);
    if (!mat) return -2;
    if (orc_build_matrix(k, k + m, mat) != 0) { free(mat); return -1; }
    const uint8_t *rows[256];
    for (int r = 0; r < m; r++) rows[r] = &mat[(k + r) * k];
    code_some_slices(rows, m, k, (const uint8_t *const *)shards, shards + k, n);
    free(mat);
    return 0;
}

/* rse/src/core.rs:640-672: recompute parity into scratch and compare */
int orc_verify(int k, int m, uint8_t *const *shards, size_t n)
{
    uint8_t *tmp[256];
    uint8_t *all[256];
    int ok = 1;
    for (int i = 0; i < k; i++) all[i] = shards[i];
    for (int r = 0; r < m; r++) {
        tmp[r] = (uint8_t *)malloc(n ? n : 1);
        all[k + r] = tmp[r];
    }
    if (orc_encode(k, m, all, n) != 0) ok = -1;
    for (int r = 0; r < m && ok == 1; r++)
        if (memcmp(tmp[r], shards[k + r], n) != 0) ok = 0;
    for (int r = 0; r < m; r++) free(tmp[r]);
    return ok;
}

/* rse/src/core.rs:700-734 + :780-830: rows of the generator for the first k present shards, inverted */
int orc_decode_matrix(int k, int m, const uint8_t *present, uint8_t *out, int *valid)
{
    int total = k + m, nv = 0;
    uint8_t *mat = (uint8_t *)malloc((size_t)total * k);
    uint8_t *sub = (uint8_t *)malloc((size_t)k * k);
    if (!mat || !sub) { free(mat); free(sub); return -2; }
    if (orc_build_matrix(k, total, mat) != 0) { free(mat); free(sub); return -1; }
    for (int r = 0; r < total && nv < k; r++)
        if (present[r]) {
            memcpy(&sub[nv * k], &mat[r * k], (size_t)k);
            valid[nv++] = r;
        }
    int rc = (nv < k) ? -3 : orc_matrix_invert(sub, k, out);
    free(mat);
    free(sub);
    return rc;
}

/* rse/src/core.rs:736-926 */
int orc_reconstruct(int k, int m, uint8_t *const *shards, const uint8_t *present, size_t n,
                    int data_only)
{
    int total = k + m, npresent = 0;
    for (int i = 0; i < total; i++) npresent += present[i] ? 1 : 0;
    if (npresent == total) return 0;   /* core.rs:766-770 */
    if (npresent < k) return -3;       /* TooFewShardsPresent core.rs:773-775 */

    uint8_t *dec = (uint8_t *)malloc((size_t)k * k);
    uint8_t *mat = (uint8_t *)malloc((size_t)total * k);
    int valid[256];
    if (!dec || !mat) { free(dec); free(mat); return -2; }
    int rc = orc_decode_matrix(k, m, present, dec, valid);
    if (rc != 0 || orc_build_matrix(k, total, mat) != 0) { free(dec); free(mat); return rc ? rc : -1; }

    const uint8_t *sub[256];
    for (int i = 0; i < k; i++) sub[i] = shards[valid[i]];

    /* missing data shards: row j of the decode matrix over the k valid shards  core.rs:852-866 */
    const uint8_t *rows[256];
    uint8_t *outs[256];
    int nout = 0;
    for (int j = 0; j < k; j++)
        if (!present[j]) { rows[nout] = &dec[j * k]; outs[nout] = shards[j]; nout++; }
    if (nout) code_some_slices(rows, nout, k, sub, outs, n);

    if (!data_only) {
        /* missing parity from the now complete data  core.rs:868-922 */
        nout = 0;
        for (int p = k; p < total; p++)
            if (!present[p]) { rows[nout] = &mat[p * k]; outs[nout] = shards[p]; nout++; }
        if (nout) code_some_slices(rows, nout, k, (const uint8_t *const *)shards, outs, n);
    }
    free(dec);
    free(mat);
    return 0;
}

/* ----------------------------------------------------------------- layout */

/* weed/storage/disk_location_ec.go:428-448 */
int64_t orc_expected_shard_size(int64_t dat_size, int k, int64_t large, int64_t small)
{
    if (k <= 0) return 0;
    int64_t large_row = large * k;
    int64_t nlarge = dat_size / large_row;
    int64_t sz = nlarge * large;
    int64_t rem = dat_size - nlarge * large_row;
    if (rem > 0) {
        int64_t small_row = small * k;
        sz += ((rem + small_row - 1) / small_row) * small;
    }
    return sz;
}

/* weed/storage/erasure_coding/ec_locate.go:55-63 */
static void next_block(int *block_index, int *is_large, int64_t nlarge_rows, int k)
{
    int nb = *block_index + 1;
    if (*is_large && (int64_t)nb == nlarge_rows * k) {
        *is_large = 0;
        nb = 0;
    }
    *block_index = nb;
}

/* weed/storage/erasure_coding/ec_locate.go:16-85 (DataShardsCount generalised to k) */
int orc_locate_data(int64_t large, int64_t small, int64_t shard_dat_size, int64_t offset,
                    int64_t size, int k, orc_interval *out, int cap)
{
    int64_t large_row = large * k;
    int64_t nlarge_rows = shard_dat_size / large;
    int is_large, block_index;
    int64_t inner;
    if (offset < nlarge_rows * large_row) {
        is_large = 1;
        block_index = (int)(offset / large);
        inner = offset % large;
    } else {
        is_large = 0;
        offset -= nlarge_rows * large_row;
        block_index = (int)(offset / small);
        inner = offset % small;
    }
    int n = 0;
    while (size > 0) {
        int64_t remaining = (is_large ? large : small) - inner;
        if (remaining <= 0) {
            next_block(&block_index, &is_large, nlarge_rows, k);
            inner = 0;
            continue;
        }
        if (n >= cap) return -1;
        orc_interval *iv = &out[n++];
        iv->block_index = block_index;
        iv->inner_block_offset = inner;
        iv->is_large_block = is_large;
        iv->large_block_rows_count = (int)nlarge_rows;
        if (size <= remaining) {
            iv->size = size;
            return n;
        }
        iv->size = remaining;
        size -= remaining;
        next_block(&block_index, &is_large, nlarge_rows, k);
        inner = 0;
    }
    return n;
}

/* weed/storage/erasure_coding/ec_locate.go:87-98 */
void orc_interval_to_shard(const orc_interval *iv, int64_t large, int64_t small, int k,
                           int *shard_id, int64_t *shard_offset)
{
    int64_t off = iv->inner_block_offset;
    int row = iv->block_index / k;
    if (iv->is_large_block) off += (int64_t)row * large;
    else off += (int64_t)iv->large_block_rows_count * large + (int64_t)row * small;
    *shard_id = iv->block_index % k;
    *shard_offset = off;
}

/* ------------------------------------------------------ in-memory file walk */

/* ec_encoder.go:248-278 on a memory image: k reads (zero-filled past EOF), Encode, k+m appends */
static int encode_one_batch(const uint8_t *dat, int64_t dat_size, int64_t start, int64_t block,
                            int k, int m, int64_t bufsz, uint8_t *const *bufs,
                            uint8_t *const *shards, int64_t *written)
{
    for (int i = 0; i < k; i++) {
        int64_t off = start + block * i;
        int64_t have = dat_size - off;
        if (have < 0) have = 0;
        if (have > bufsz) have = bufsz;
        if (have > 0) memcpy(bufs[i], dat + off, (size_t)have);
        if (have < bufsz) memset(bufs[i] + have, 0, (size_t)(bufsz - have));
    }
    int rc = orc_encode(k, m, bufs, (size_t)bufsz);
    if (rc) return rc;
    for (int i = 0; i < k + m; i++) memcpy(shards[i] + *written, bufs[i], (size_t)bufsz);
    *written += bufsz;
    return 0;
}

/* ec_encoder.go:202-222 + :280-321 */
int orc_encode_dat_image(const uint8_t *dat, int64_t dat_size, int k, int m, int64_t bufsz,
                         int64_t large, int64_t small, uint8_t *const *shards)
{
    if (bufsz <= 0 || large % bufsz || small % bufsz) return -1; /* glog.Fatalf ec_encoder.go:210-212 */
    uint8_t *bufs[ORC_MAX_SHARDS];
    for (int i = 0; i < k + m; i++) bufs[i] = (uint8_t *)malloc((size_t)bufsz);
    int64_t remaining = dat_size, processed = 0, written = 0;
    int64_t large_row = large * k, small_row = small * k;
    int rc = 0;
    while (rc == 0 && remaining >= large_row) {           /* ec_encoder.go:304-311 */
        for (int64_t b = 0; rc == 0 && b < large / bufsz; b++)
            rc = encode_one_batch(dat, dat_size, processed + b * bufsz, large, k, m, bufsz, bufs, shards, &written);
        remaining -= large_row;
        processed += large_row;
    }
    while (rc == 0 && remaining > 0) {                    /* ec_encoder.go:312-319 */
        for (int64_t b = 0; rc == 0 && b < small / bufsz; b++)
            rc = encode_one_batch(dat, dat_size, processed + b * bufsz, small, k, m, bufsz, bufs, shards, &written);
        remaining -= small_row;
        processed += small_row;
    }
    for (int i = 0; i < k + m; i++) free(bufs[i]);
    return rc;
}

/* ec_decoder.go:176-223 */
int orc_write_dat_image(uint8_t *dat, int64_t dat_size, int k, int64_t large, int64_t small,
                        const uint8_t *const *shards)
{
    int64_t pos[ORC_MAX_SHARDS] = {0};
    int64_t out = 0, remaining = dat_size;
    while (remaining >= (int64_t)k * large) {
        for (int s = 0; s < k; s++) {
            memcpy(dat + out, shards[s] + pos[s], (size_t)large);
            pos[s] += large;
            out += large;
            remaining -= large;
        }
    }
    while (remaining > 0) {
        for (int s = 0; s < k; s++) {
            int64_t n = remaining < small ? remaining : small;
            if (n > 0) memcpy(dat + out, shards[s] + pos[s], (size_t)n);
            pos[s] += n;
            out += n;
            remaining -= n;
        }
    }
    return 0;
}

/* --------------------------------------------------------------- file level */

static int ext_name(char *dst, size_t cap, const char *base, int idx)
{
    return snprintf(dst, cap, "%s.ec%02d", base, idx) < (int)cap ? 0 : -1; /* ec_encoder.go:106-108 */
}

/* ec_encoder.go:110-128 → encodeDatFile :280-321, streaming through pread/write like the reference */
int orc_generate_ec_files(const char *base, int64_t bufsz, int64_t large, int64_t small, int k, int m)
{
    char path[4096];
    if (snprintf(path, sizeof path, "%s.dat", base) >= (int)sizeof path) return -ENAMETOOLONG;
    int fd = open(path, O_RDONLY);
    if (fd < 0) return -errno;
    struct stat st;
    if (fstat(fd, &st) != 0) { int e = errno; close(fd); return -e; }
    if (bufsz <= 0 || large % bufsz || small % bufsz) { close(fd); return -EINVAL; }

    int total = k + m, rc = 0;
    int outs[ORC_MAX_SHARDS];
    uint8_t *bufs[ORC_MAX_SHARDS];
    for (int i = 0; i < total; i++) { outs[i] = -1; bufs[i] = NULL; }
    for (int i = 0; i < total && rc == 0; i++) {          /* openEcFiles ec_encoder.go:224-238 */
        if (ext_name(path, sizeof path, base, i)) { rc = -ENAMETOOLONG; break; }
        outs[i] = open(path, O_TRUNC | O_CREAT | O_WRONLY, 0644);
        if (outs[i] < 0) rc = -errno;
        bufs[i] = (uint8_t *)malloc((size_t)bufsz);
        if (!bufs[i]) rc = -ENOMEM;
    }
    int64_t remaining = st.st_size, processed = 0;
    int64_t large_row = large * k, small_row = small * k;
    for (int pass = 0; pass < 2 && rc == 0; pass++) {
        int64_t block = pass == 0 ? large : small;
        int64_t row = pass == 0 ? large_row : small_row;
        while (rc == 0 && (pass == 0 ? remaining >= row : remaining > 0)) {
            for (int64_t b = 0; rc == 0 && b < block / bufsz; b++) {
                int64_t start = processed + b * bufsz;
                for (int i = 0; i < k; i++) {              /* ec_encoder.go:251-263 */
                    ssize_t got = pread(fd, bufs[i], (size_t)bufsz, (off_t)(start + block * i));
                    if (got < 0) { rc = -errno; break; }
                    if (got < bufsz) memset(bufs[i] + got, 0, (size_t)(bufsz - got));
                }
                if (rc == 0) rc = orc_encode(k, m, bufs, (size_t)bufsz);
                for (int i = 0; i < total && rc == 0; i++)   /* ec_encoder.go:270-275 */
                    if (write(outs[i], bufs[i], (size_t)bufsz) != bufsz) rc = -EIO;
            }
            remaining -= row;
            processed += row;
        }
    }
    for (int i = 0; i < total; i++) {
        if (outs[i] >= 0) close(outs[i]);
        free(bufs[i]);
    }
    close(fd);
    return rc;
}

/* ec_encoder.go:146-200 + :323-377 (base directory only; additionalDirs is host logic, tested in the product) */
int orc_rebuild_ec_files(const char *base, int k, int m, uint32_t *rebuilt, int *nrebuilt)
{
    enum { BUF = 1024 * 1024 }; /* ErasureCodingSmallBlockSize ec_encoder.go:333 */
    int total = k + m, npresent = 0, rc = 0;
    int in[ORC_MAX_SHARDS], out[ORC_MAX_SHARDS];
    uint8_t present[ORC_MAX_SHARDS];
    uint8_t *bufs[ORC_MAX_SHARDS];
    char path[4096];
    *nrebuilt = 0;
    for (int i = 0; i < total; i++) { in[i] = out[i] = -1; bufs[i] = NULL; }
    for (int i = 0; i < total; i++) {                     /* pass 1 :150-169 */
        if (ext_name(path, sizeof path, base, i)) return -ENAMETOOLONG;
        in[i] = open(path, O_RDONLY);
        present[i] = in[i] >= 0;
        if (present[i]) npresent++;
        else rebuilt[(*nrebuilt)++] = (uint32_t)i;
    }
    if (npresent < k) rc = -ENODATA;                      /* :172-175 before any output is created */
    for (int i = 0; i < total && rc == 0; i++) {          /* pass 2 :182-193 */
        bufs[i] = (uint8_t *)malloc(BUF);
        if (present[i]) continue;
        ext_name(path, sizeof path, base, i);
        out[i] = open(path, O_TRUNC | O_WRONLY | O_CREAT, 0644);
        if (out[i] < 0) rc = -errno;
    }
    int64_t start = 0;
    ssize_t chunk = 0;
    while (rc == 0) {                                     /* :340-376 */
        int done = 0;
        for (int i = 0; i < total; i++) {
            if (!present[i]) continue;
            ssize_t got = pread(in[i], bufs[i], BUF, (off_t)start);
            if (got <= 0) { done = 1; break; }
            if (chunk == 0) chunk = got;
            if (chunk != got) { rc = -EPROTO; break; }    /* "ec shard size expected %d actual %d" :351-353 */
        }
        if (done || rc) break;
        rc = orc_reconstruct(k, m, bufs, present, BUF, 0);
        for (int i = 0; i < total && rc == 0; i++)
            if (!present[i] && pwrite(out[i], bufs[i], (size_t)chunk, (off_t)start) != chunk) rc = -EIO;
        start += chunk;
    }
    for (int i = 0; i < total; i++) {
        if (in[i] >= 0) close(in[i]);
        if (out[i] >= 0) close(out[i]);
        free(bufs[i]);
    }
    return rc;
}

/* ---------------------------------------------------------- synthetic data */

static inline uint64_t splitmix64_at(uint64_t seed, uint64_t j)
{
    uint64_t z = seed + (j + 1) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

/* byte b of the stream = byte (b%8) (little-endian) of splitmix64_at(seed, b/8)  — SURVEY §8(d) */
void orc_synth_fill(uint8_t *dst, uint64_t byte_offset, size_t n, uint64_t seed)
{
    for (size_t i = 0; i < n; i++) {
        uint64_t b = byte_offset + i;
        dst[i] = (uint8_t)(splitmix64_at(seed, b >> 3) >> ((b & 7) * 8));
    }
}
