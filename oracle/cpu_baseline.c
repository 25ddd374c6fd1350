/*
 * oracle/cpu_baseline.c — CPU BASELINE DRIVER (test/bench infrastructure, NOT product code).
 *
 * Times the reference's CPU arithmetic for the RS(k,m) encode / reconstruct path on host cores:
 *
 *   kind 0 "reference": the reference's own vendored SIMD kernel, compiled unmodified from
 *          /root/reference/seaweed-volume/vendor/reed-solomon-erasure/simd_c/reedsolomon.c into
 *          oracle/_ref/libref_rs_<isa>.so and driven exactly like code_some_slices
 *          (rse/src/core.rs:484-512: for each input, for each output row: first input mul, the
 *          rest mul_xor), in 256 KiB batches like encodeDataOneBatch
 *          (weed/storage/erasure_coding/ec_encoder.go:68,248-278).
 *   kind 1 "gfni port": a fused AVX-512/AVX2 + GFNI kernel (vgf2p8affineqb, all k inputs of a
 *          64-byte column combined in registers into the m outputs) — a restatement of what
 *          klauspost/reedsolomon v1.14.0 (go.mod:50, source not in the tree) runs on a GFNI CPU
 *          (its mulGFNI_10x4_64 family).  Same matrix, same field; labelled "port".
 *
 * Threads split the byte-column range (the way klauspost splits an Encode across goroutines).
 * The matrix rows are passed in, so the same driver serves encode (parity rows) and reconstruct
 * (decode rows).
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <immintrin.h>
#include <pthread.h>
#include <sched.h>
#include <time.h>
#include <unistd.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "rs_oracle.h"

typedef size_t (*gal_fn)(const uint8_t low[16], const uint8_t high[16], const uint8_t *in,
                         uint8_t *out, size_t len); /* rse/simd_c/reedsolomon.h:30-42 */

static void *g_ref_handle;
static gal_fn g_ref_mul, g_ref_mul_xor;
static char g_ref_isa[32] = "none";

/* Load oracle/_ref/libref_rs_<isa>.so for the best ISA this CPU supports. dir = oracle/_ref */
int orc_ref_load(const char *dir)
{
    if (g_ref_handle) return 0;
    __builtin_cpu_init();
    const char *order[3];
    int n = 0;
    if (__builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw")) order[n++] = "avx512";
    if (__builtin_cpu_supports("avx2")) order[n++] = "avx2";
    if (__builtin_cpu_supports("ssse3")) order[n++] = "ssse3";
    for (int i = 0; i < n; i++) {
        char path[4096];
        snprintf(path, sizeof path, "%s/libref_rs_%s.so", dir, order[i]);
        void *h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
        if (!h) continue;
        g_ref_mul = (gal_fn)dlsym(h, "reedsolomon_gal_mul");
        g_ref_mul_xor = (gal_fn)dlsym(h, "reedsolomon_gal_mul_xor");
        if (g_ref_mul && g_ref_mul_xor) {
            g_ref_handle = h;
            snprintf(g_ref_isa, sizeof g_ref_isa, "%s", order[i]);
            return 0;
        }
        dlclose(h);
    }
    return -1;
}

const char *orc_ref_isa(void) { return g_ref_isa; }

int orc_cpu_has_gfni(void)
{
    __builtin_cpu_init();
    if (!__builtin_cpu_supports("gfni")) return 0;
    if (__builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw")) return 2;
    if (__builtin_cpu_supports("avx2")) return 1;
    return 0;
}

/* The reference kernel handles whole vectors only and returns bytes done; the caller finishes the
 * tail with the scalar loop (rse/src/galois_8.rs:291-327). */
static void ref_mul(uint8_t c, const uint8_t *in, uint8_t *out, size_t n, int xor_into)
{
    uint8_t low[16], high[16];
    orc_mul_table_half(c, low, high);
    size_t done = xor_into ? g_ref_mul_xor(low, high, in, out, n) : g_ref_mul(low, high, in, out, n);
    if (done < n) {
        if (xor_into) orc_mul_slice_xor(c, in + done, out + done, n - done);
        else orc_mul_slice(c, in + done, out + done, n - done);
    }
}

/* 8×8 bit matrix A with gf2p8affineqb(x, A, 0) == c ⊗ x over 0x11D.
 * Output bit i of each byte = parity(A.byte[7-i] & x) (Intel SDM, GF2P8AFFINEQB). */
static uint64_t gfni_matrix(uint8_t c)
{
    uint64_t a = 0;
    for (int i = 0; i < 8; i++) {
        uint8_t row = 0;
        for (int j = 0; j < 8; j++)
            if ((orc_mul(c, (uint8_t)(1u << j)) >> i) & 1) row |= (uint8_t)(1u << j);
        a |= (uint64_t)row << (8 * (7 - i));
    }
    return a;
}

__attribute__((target("avx512f,avx512bw,gfni")))
static void gfni512_range(int k, int r, const uint64_t *mats /* r*k */, const uint8_t *const *in,
                          uint8_t *const *out, size_t lo, size_t hi)
{
    size_t x = lo;
    for (; x + 64 <= hi; x += 64) {
        __m512i acc[ORC_MAX_SHARDS];
        for (int i = 0; i < k; i++) {
            __m512i v = _mm512_loadu_si512((const void *)(in[i] + x));
            for (int p = 0; p < r; p++) {
                __m512i t = _mm512_gf2p8affine_epi64_epi8(v, _mm512_set1_epi64((long long)mats[p * k + i]), 0);
                acc[p] = i == 0 ? t : _mm512_xor_si512(acc[p], t);
            }
        }
        for (int p = 0; p < r; p++) _mm512_storeu_si512((void *)(out[p] + x), acc[p]);
    }
    for (; x < hi; x++)
        for (int p = 0; p < r; p++) {
            uint8_t v = 0;
            for (int i = 0; i < k; i++) {
                /* scalar tail through the same bit matrices */
                uint64_t a = mats[p * k + i];
                uint8_t b = in[i][x], o = 0;
                for (int bit = 0; bit < 8; bit++)
                    o |= (uint8_t)(__builtin_parity((unsigned)((a >> (8 * (7 - bit))) & 0xff) & b) << bit);
                v ^= o;
            }
            out[p][x] = v;
        }
}

/* RS(10,4)-shaped fast path: 4 accumulators live in registers across the 10 inputs. */
__attribute__((target("avx512f,avx512bw,gfni")))
static void gfni512_range_r4(int k, const uint64_t *mats, const uint8_t *const *in,
                             uint8_t *const *out, size_t lo, size_t hi)
{
    size_t x = lo;
    for (; x + 64 <= hi; x += 64) {
        __m512i v = _mm512_loadu_si512((const void *)(in[0] + x));
        __m512i a0 = _mm512_gf2p8affine_epi64_epi8(v, _mm512_set1_epi64((long long)mats[0 * k]), 0);
        __m512i a1 = _mm512_gf2p8affine_epi64_epi8(v, _mm512_set1_epi64((long long)mats[1 * k]), 0);
        __m512i a2 = _mm512_gf2p8affine_epi64_epi8(v, _mm512_set1_epi64((long long)mats[2 * k]), 0);
        __m512i a3 = _mm512_gf2p8affine_epi64_epi8(v, _mm512_set1_epi64((long long)mats[3 * k]), 0);
        for (int i = 1; i < k; i++) {
            v = _mm512_loadu_si512((const void *)(in[i] + x));
            a0 = _mm512_xor_si512(a0, _mm512_gf2p8affine_epi64_epi8(v, _mm512_set1_epi64((long long)mats[0 * k + i]), 0));
            a1 = _mm512_xor_si512(a1, _mm512_gf2p8affine_epi64_epi8(v, _mm512_set1_epi64((long long)mats[1 * k + i]), 0));
            a2 = _mm512_xor_si512(a2, _mm512_gf2p8affine_epi64_epi8(v, _mm512_set1_epi64((long long)mats[2 * k + i]), 0));
            a3 = _mm512_xor_si512(a3, _mm512_gf2p8affine_epi64_epi8(v, _mm512_set1_epi64((long long)mats[3 * k + i]), 0));
        }
        _mm512_stream_si512((void *)(out[0] + x), a0);
        _mm512_stream_si512((void *)(out[1] + x), a1);
        _mm512_stream_si512((void *)(out[2] + x), a2);
        _mm512_stream_si512((void *)(out[3] + x), a3);
    }
    if (x < hi) gfni512_range(k, 4, mats, in, out, x, hi);
}

__attribute__((target("avx2,gfni")))
static void gfni256_range(int k, int r, const uint64_t *mats, const uint8_t *const *in,
                          uint8_t *const *out, size_t lo, size_t hi)
{
    size_t x = lo;
    for (; x + 32 <= hi; x += 32) {
        __m256i acc[ORC_MAX_SHARDS];
        for (int i = 0; i < k; i++) {
            __m256i v = _mm256_loadu_si256((const __m256i *)(in[i] + x));
            for (int p = 0; p < r; p++) {
                __m256i t = _mm256_gf2p8affine_epi64_epi8(v, _mm256_set1_epi64x((long long)mats[p * k + i]), 0);
                acc[p] = i == 0 ? t : _mm256_xor_si256(acc[p], t);
            }
        }
        for (int p = 0; p < r; p++) _mm256_storeu_si256((__m256i *)(out[p] + x), acc[p]);
    }
    for (; x < hi; x++)
        for (int p = 0; p < r; p++) {
            uint8_t v = 0;
            for (int i = 0; i < k; i++) {
                uint64_t a = mats[p * k + i];
                uint8_t b = in[i][x], o = 0;
                for (int bit = 0; bit < 8; bit++)
                    o |= (uint8_t)(__builtin_parity((unsigned)((a >> (8 * (7 - bit))) & 0xff) & b) << bit);
                v ^= o;
            }
            out[p][x] = v;
        }
}

typedef struct {
    int kind, k, r;
    const uint8_t *rows;  /* r*k coefficients */
    const uint64_t *mats; /* r*k gfni matrices (kind 1) */
    const uint8_t *const *in;
    uint8_t *const *out;
    size_t lo, hi, batch;
} job_t;

static void *worker(void *arg)
{
    job_t *j = (job_t *)arg;
    if (j->kind == 0) {
        /* code_some_slices over batches of `batch` bytes (256 KiB in production) */
        for (size_t b = j->lo; b < j->hi; b += j->batch) {
            size_t n = j->hi - b < j->batch ? j->hi - b : j->batch;
            for (int i = 0; i < j->k; i++)
                for (int p = 0; p < j->r; p++)
                    ref_mul(j->rows[p * j->k + i], j->in[i] + b, j->out[p] + b, n, i != 0);
        }
    } else {
        int lvl = orc_cpu_has_gfni();
        if (lvl == 2) {
            if (j->r == 4 && (((uintptr_t)j->out[0] | (uintptr_t)j->out[1] | (uintptr_t)j->out[2] |
                               (uintptr_t)j->out[3] | j->lo) & 63) == 0)
                gfni512_range_r4(j->k, j->mats, j->in, j->out, j->lo, j->hi);
            else
                gfni512_range(j->k, j->r, j->mats, j->in, j->out, j->lo, j->hi);
        } else {
            gfni256_range(j->k, j->r, j->mats, j->in, j->out, j->lo, j->hi);
        }
    }
    return NULL;
}

/*
 * out[p][x] = XOR_i rows[p*k+i] ⊗ in[i][x] for x in [0,n), split over `threads` threads.
 * kind 0 = reference kernel (needs orc_ref_load), kind 1 = gfni port.  0 ok; -1 unavailable.
 */
int orc_cpu_apply_mt(int kind, int k, int r, const uint8_t *rows, const uint8_t *const *in,
                     uint8_t *const *out, size_t n, int threads, size_t batch)
{
    if (kind == 0 && !g_ref_handle) return -1;
    if (kind == 1 && !orc_cpu_has_gfni()) return -1;
    if (threads < 1) threads = 1;
    if (batch == 0) batch = 256 * 1024;
    uint64_t *mats = NULL;
    if (kind == 1) {
        mats = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)r * k);
        for (int i = 0; i < r * k; i++) mats[i] = gfni_matrix(rows[i]);
    }
    /* per-thread ranges in multiples of 4 KiB so streaming stores stay aligned */
    size_t unit = 4096, units = (n + unit - 1) / unit;
    job_t *jobs = (job_t *)calloc((size_t)threads, sizeof(job_t));
    pthread_t *tids = (pthread_t *)calloc((size_t)threads, sizeof(pthread_t));
    int started = 0;
    for (int t = 0; t < threads; t++) {
        size_t lo = units * t / threads * unit, hi = units * (t + 1) / threads * unit;
        if (hi > n) hi = n;
        if (lo >= hi) continue;
        jobs[started] = (job_t){kind, k, r, rows, mats, in, out, lo, hi, batch};
        if (threads == 1) worker(&jobs[started]);
        else pthread_create(&tids[started], NULL, worker, &jobs[started]);
        started++;
    }
    if (threads > 1)
        for (int t = 0; t < started; t++) pthread_join(tids[t], NULL);
    free(jobs);
    free(tids);
    free(mats);
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * Whole-box throughput bench: every thread owns its buffers (allocated and first-touched by the
 * thread itself, so they are NUMA-local), threads are pinned, started together on a barrier and each
 * runs `passes` passes over its own k inputs → r outputs of `bytes` each.  Returns input GB/s
 * (1e9) = threads·passes·k·bytes / wall seconds, or a negative value if the kind is unavailable.
 * This is the CPU's best case: no thread start-up in the timed region, no remote-socket traffic. */
typedef struct {
    int kind, k, r, id;
    const uint8_t *rows;
    const uint64_t *mats;
    size_t bytes, batch;
    int passes;
    pthread_barrier_t *start, *stop;
    int ok;
} bench_job_t;

static void *bench_worker(void *arg)
{
    bench_job_t *j = (bench_job_t *)arg;
    cpu_set_t set;
    long ncpu = sysconf(_SC_NPROCESSORS_ONLN);
    if (ncpu > 0) {
        CPU_ZERO(&set);
        CPU_SET((int)(j->id % ncpu), &set);
        pthread_setaffinity_np(pthread_self(), sizeof set, &set); /* best effort */
    }
    uint8_t *in[ORC_MAX_SHARDS], *out[ORC_MAX_SHARDS];
    j->ok = 1;
    for (int i = 0; i < j->k; i++) {
        if (posix_memalign((void **)&in[i], 4096, j->bytes)) { j->ok = 0; in[i] = NULL; continue; }
        uint64_t s = 0x9E3779B97F4A7C15ull * (uint64_t)(j->id * 64 + i + 1);
        for (size_t x = 0; x + 8 <= j->bytes; x += 8) { /* xorshift fill = first touch */
            s ^= s << 13; s ^= s >> 7; s ^= s << 17;
            memcpy(in[i] + x, &s, 8);
        }
    }
    for (int p = 0; p < j->r; p++) {
        if (posix_memalign((void **)&out[p], 4096, j->bytes)) { j->ok = 0; out[p] = NULL; continue; }
        memset(out[p], 0, j->bytes);
    }
    job_t w = {j->kind, j->k, j->r, j->rows, j->mats, (const uint8_t *const *)in, out, 0, j->bytes, j->batch};
    if (j->ok) worker(&w); /* warm-up pass */
    pthread_barrier_wait(j->start);
    if (j->ok)
        for (int p = 0; p < j->passes; p++) worker(&w);
    pthread_barrier_wait(j->stop);
    for (int i = 0; i < j->k; i++) free(in[i]);
    for (int p = 0; p < j->r; p++) free(out[p]);
    return NULL;
}

double orc_cpu_bench(int kind, int k, int r, const uint8_t *rows, size_t bytes, int threads, int passes,
                     size_t batch)
{
    if (kind == 0 && !g_ref_handle) return -1.0;
    if (kind == 1 && !orc_cpu_has_gfni()) return -1.0;
    if (threads < 1) threads = 1;
    if (batch == 0) batch = 256 * 1024;
    uint64_t *mats = NULL;
    if (kind == 1) {
        mats = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)r * k);
        for (int i = 0; i < r * k; i++) mats[i] = gfni_matrix(rows[i]);
    }
    pthread_barrier_t start, stop;
    pthread_barrier_init(&start, NULL, (unsigned)threads + 1);
    pthread_barrier_init(&stop, NULL, (unsigned)threads + 1);
    bench_job_t *jobs = (bench_job_t *)calloc((size_t)threads, sizeof(bench_job_t));
    pthread_t *tids = (pthread_t *)calloc((size_t)threads, sizeof(pthread_t));
    for (int t = 0; t < threads; t++) {
        jobs[t] = (bench_job_t){kind, k, r, t, rows, mats, bytes, batch, passes, &start, &stop, 0};
        pthread_create(&tids[t], NULL, bench_worker, &jobs[t]);
    }
    struct timespec t0, t1;
    pthread_barrier_wait(&start);
    clock_gettime(CLOCK_MONOTONIC, &t0);
    pthread_barrier_wait(&stop);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    int ok = 1;
    for (int t = 0; t < threads; t++) { pthread_join(tids[t], NULL); ok &= jobs[t].ok; }
    pthread_barrier_destroy(&start);
    pthread_barrier_destroy(&stop);
    free(jobs);
    free(tids);
    free(mats);
    if (!ok) return -2.0;
    double secs = (double)(t1.tv_sec - t0.tv_sec) + (double)(t1.tv_nsec - t0.tv_nsec) * 1e-9;
    return (double)threads * passes * k * (double)bytes / secs / 1e9;
}

/* ------------------------------------------------------------------------------------------------
 * B3 of BASELINE.md: the reference-shaped file walk at SIMD speed — generateEcFiles exactly as
 * ec_encoder.go:110-128,202-321 does it (one thread; per 256 KiB batch: 10 × pread, Encode,
 * 14 × write, strictly serial), with the Encode done by the reference's own compiled kernel
 * (kind 0) or the GFNI port (kind 1).  Returns 0 or a negative errno. */
#include <errno.h>
#include <fcntl.h>
#include <sys/stat.h>

int orc_generate_ec_files_simd(const char *base, int kind, int64_t bufsz, int64_t large, int64_t small, int k, int m)
{
    if (kind == 0 && !g_ref_handle) return -ENOSYS;
    if (kind == 1 && !orc_cpu_has_gfni()) return -ENOSYS;
    if (bufsz <= 0 || large % bufsz || small % bufsz || k + m > ORC_MAX_SHARDS) return -EINVAL;
    char path[4096];
    snprintf(path, sizeof path, "%s.dat", base);
    int fd = open(path, O_RDONLY);
    if (fd < 0) return -errno;
    struct stat st;
    fstat(fd, &st);
    uint8_t *gen = (uint8_t *)malloc((size_t)(k + m) * k);
    orc_build_matrix(k, k + m, gen);
    const uint8_t *rows = gen + (size_t)k * k;
    int outs[ORC_MAX_SHARDS];
    uint8_t *bufs[ORC_MAX_SHARDS];
    int rc = 0;
    for (int i = 0; i < k + m; i++) {
        snprintf(path, sizeof path, "%s.ec%02d", base, i);
        outs[i] = open(path, O_TRUNC | O_CREAT | O_WRONLY, 0644);
        if (outs[i] < 0) rc = -errno;
        if (posix_memalign((void **)&bufs[i], 4096, (size_t)bufsz)) rc = -ENOMEM;
    }
    int64_t remaining = st.st_size, processed = 0;
    for (int pass = 0; pass < 2 && rc == 0; pass++) {
        int64_t block = pass == 0 ? large : small, row = block * k;
        while (rc == 0 && (pass == 0 ? remaining >= row : remaining > 0)) {
            for (int64_t b = 0; rc == 0 && b < block / bufsz; b++) {
                for (int i = 0; i < k; i++) {
                    ssize_t got = pread(fd, bufs[i], (size_t)bufsz, (off_t)(processed + b * bufsz + block * i));
                    if (got < 0) { rc = -errno; break; }
                    if (got < bufsz) memset(bufs[i] + got, 0, (size_t)(bufsz - got));
                }
                if (rc == 0)
                    rc = orc_cpu_apply_mt(kind, k, m, rows, (const uint8_t *const *)bufs, bufs + k, (size_t)bufsz, 1, (size_t)bufsz);
                for (int i = 0; i < k + m && rc == 0; i++)
                    if (write(outs[i], bufs[i], (size_t)bufsz) != bufsz) rc = -EIO;
            }
            remaining -= row;
            processed += row;
        }
    }
    for (int i = 0; i < k + m; i++) {
        if (outs[i] >= 0) close(outs[i]);
        free(bufs[i]);
    }
    free(gen);
    close(fd);
    return rc;
}

/* ------------------------------------------------------------------------------------------------
 * Whole-volume expected digests (CHECKER for the full-size GPU runs; test/bench infrastructure).
 *
 * The synthetic volume of the GPU bench is byte b = splitmix64 stream of (seed, b) (orc_synth_fill).
 * This walks the shard space of encodeDatFile (ec_encoder.go:280-321: rows of k large blocks while
 * remaining >= k*large, then rows of k small blocks, bytes past EOF read as zero, :258-262) in
 * chunks, regenerates the k data columns of each chunk straight from the generator, encodes them
 * with the reference's arithmetic (kind 0 = reference C kernel from oracle/_ref, 1 = GFNI port,
 * 2 = the scalar restatement orc_encode's tables) and folds every shard into the same 64-bit
 * order-sensitive digest the device computes (Σ_j splitmix64_at(word_j, j) mod 2^64 over the
 * shard's little-endian 8-byte words).  Nothing of the volume is ever held in memory, so a 30 GiB
 * volume costs a few seconds on a dozen cores and can be checked in full after every GPU run.
 *   digests[0..k)   data shards (.ec00 … ), digests[k..k+m) parity shards.
 * Returns 0, -1 (kind unavailable) or -2 (bad arguments). */
static inline uint64_t vd_mix(uint64_t seed, uint64_t j)
{
    uint64_t z = seed + (j + 1) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

typedef struct {
    int kind, k, m, id, threads;
    int64_t dat_size, large, small, shard_size, chunk;
    uint64_t seed;
    const uint8_t *rows;
    const uint64_t *mats;
    uint64_t sums[ORC_MAX_SHARDS];
    int rc;
} vd_job_t;

static void *vd_worker(void *arg)
{
    vd_job_t *j = (vd_job_t *)arg;
    const int k = j->k, m = j->m;
    const int64_t c = j->chunk;
    uint8_t *buf[ORC_MAX_SHARDS];
    memset(j->sums, 0, sizeof j->sums);
    for (int i = 0; i < k + m; i++)
        if (posix_memalign((void **)&buf[i], 4096, (size_t)c)) { j->rc = -2; return NULL; }
    const int64_t large_row = j->large * k, small_row = j->small * k;
    const int64_t nlarge = j->dat_size / large_row;
    const int64_t nchunks = j->shard_size / c;
    for (int64_t ci = j->id; ci < nchunks; ci += j->threads) {
        const int64_t off = ci * c; /* offset inside every shard file */
        for (int i = 0; i < k; i++) {
            int64_t src; /* .dat offset of shard i's bytes [off, off+c) */
            if (off < nlarge * j->large) src = (off / j->large) * large_row + i * j->large + off % j->large;
            else {
                const int64_t o2 = off - nlarge * j->large;
                src = nlarge * large_row + (o2 / j->small) * small_row + i * j->small + o2 % j->small;
            }
            uint64_t *w = (uint64_t *)buf[i];
            const int64_t have = j->dat_size - src; /* bytes of this run that exist in the .dat */
            const int64_t nw = c / 8;
            for (int64_t x = 0; x < nw; x++) {
                const int64_t b = 8 * x;
                uint64_t v = 0;
                if (b < have) {
                    v = vd_mix(j->seed, (uint64_t)(src + b) >> 3);
                    if (b + 8 > have) v &= (~0ull) >> (8 * (8 - (have - b))); /* ragged EOF inside a word */
                }
                w[x] = v;
            }
        }
        if (j->kind == 2) {
            const uint8_t *mt = orc_mul_table();
            for (int p = 0; p < m; p++) {
                memset(buf[k + p], 0, (size_t)c);
                for (int i = 0; i < k; i++) {
                    const uint8_t *row = mt + 256 * (size_t)j->rows[p * k + i];
                    for (int64_t x = 0; x < c; x++) buf[k + p][x] ^= row[buf[i][x]];
                }
            }
        } else {
            job_t w = {j->kind, k, m, j->rows, j->mats, (const uint8_t *const *)buf, buf + k, 0, (size_t)c, 256 * 1024};
            worker(&w);
        }
        for (int i = 0; i < k + m; i++) {
            const uint64_t *w = (const uint64_t *)buf[i];
            const uint64_t j0 = (uint64_t)off >> 3;
            uint64_t s = 0;
            for (int64_t x = 0; x < c / 8; x++) s += vd_mix(w[x], j0 + (uint64_t)x);
            j->sums[i] += s;
        }
    }
    for (int i = 0; i < k + m; i++) free(buf[i]);
    return NULL;
}

int orc_volume_digests(int kind, int64_t dat_size, uint64_t seed, int k, int m, int64_t large, int64_t small,
                       int threads, uint64_t *digests)
{
    if (kind == 0 && !g_ref_handle) return -1;
    if (kind == 1 && !orc_cpu_has_gfni()) return -1;
    if (k <= 0 || m <= 0 || k + m > ORC_MAX_SHARDS || dat_size < 0 || large <= 0 || small <= 0 || !digests ||
        large % small || small % 8 || (kind != 0 && kind != 1 && kind != 2))
        return -2;
    if (threads < 1) threads = 1;
    const int64_t shard_size = orc_expected_shard_size(dat_size, k, large, small);
    int64_t chunk = small; /* divides every block; 64 KiB..1 MiB keeps a thread's 14 columns in L2 */
    while (chunk > (1 << 18) && chunk % 2 == 0 && (chunk / 2) % 64 == 0) chunk /= 2;
    uint8_t *gen = (uint8_t *)malloc((size_t)(k + m) * k);
    orc_build_matrix(k, k + m, gen);
    const uint8_t *rows = gen + (size_t)k * k;
    uint64_t *mats = NULL;
    if (kind == 1) {
        mats = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)m * k);
        for (int i = 0; i < m * k; i++) mats[i] = gfni_matrix(rows[i]);
    }
    vd_job_t *jobs = (vd_job_t *)calloc((size_t)threads, sizeof(vd_job_t));
    pthread_t *tids = (pthread_t *)calloc((size_t)threads, sizeof(pthread_t));
    for (int t = 0; t < threads; t++) {
        jobs[t] = (vd_job_t){.kind = kind, .k = k, .m = m, .id = t, .threads = threads, .dat_size = dat_size,
                             .large = large, .small = small, .shard_size = shard_size, .chunk = chunk,
                             .seed = seed, .rows = rows, .mats = mats};
        pthread_create(&tids[t], NULL, vd_worker, &jobs[t]);
    }
    int rc = 0;
    memset(digests, 0, sizeof(uint64_t) * (size_t)(k + m));
    for (int t = 0; t < threads; t++) {
        pthread_join(tids[t], NULL);
        if (jobs[t].rc) rc = jobs[t].rc;
        for (int i = 0; i < k + m; i++) digests[i] += jobs[t].sums[i];
    }
    free(jobs);
    free(tids);
    free(mats);
    free(gen);
    return rc;
}

/* ------------------------------------------------------------------------------------------------
 * The CPU arm of the file-level comparison at ITS best schedule (bench infrastructure, not product): generateEcFiles
 * with the same layout and the same shard bytes as ec_encoder.go:202-321, but scheduled the way the GPU pipeline is —
 * many stripes in flight, large contiguous I/O, reads ∥ arithmetic ∥ writes — instead of the reference's serial
 * 256 KiB loop.  Every thread owns every `threads`-th chunk of the shard space: k preads of its chunk of the .dat
 * (layout offsets, zero past EOF), the SIMD encode (kind 0 = reference C kernel, 1 = GFNI port) on buffers that
 * stay in its L2, k+m pwrites at the chunk's shard offset.  No cross-thread synchronisation at all.
 * Answers "is the GPU file path fast because of the GPU or because of the schedule?".  0 or -errno. */
typedef struct {
    int kind, k, m, id, threads, dat, *outs;
    int64_t dat_size, large, small, shard_size, chunk;
    const uint8_t *rows;
    const uint64_t *mats;
    int rc;
} gm_job_t;

static void *gm_worker(void *arg)
{
    gm_job_t *j = (gm_job_t *)arg;
    const int k = j->k, m = j->m;
    const int64_t c = j->chunk;
    uint8_t *buf[ORC_MAX_SHARDS];
    for (int i = 0; i < k + m; i++)
        if (posix_memalign((void **)&buf[i], 4096, (size_t)c)) { j->rc = -ENOMEM; return NULL; }
    const int64_t large_row = j->large * k, small_row = j->small * k;
    const int64_t nlarge = j->dat_size / large_row;
    const int64_t nchunks = j->shard_size / c;
    for (int64_t ci = j->id; ci < nchunks && j->rc == 0; ci += j->threads) {
        const int64_t off = ci * c;
        for (int i = 0; i < k; i++) {
            int64_t src;
            if (off < nlarge * j->large) src = (off / j->large) * large_row + i * j->large + off % j->large;
            else {
                const int64_t o2 = off - nlarge * j->large;
                src = nlarge * large_row + (o2 / j->small) * small_row + i * j->small + o2 % j->small;
            }
            int64_t got = 0;
            while (got < c) {
                ssize_t n = pread(j->dat, buf[i] + got, (size_t)(c - got), (off_t)(src + got));
                if (n < 0) { if (errno == EINTR) continue; j->rc = -errno; break; }
                if (n == 0) break;
                got += n;
            }
            if (got < c) memset(buf[i] + got, 0, (size_t)(c - got));
        }
        job_t w = {j->kind, k, m, j->rows, j->mats, (const uint8_t *const *)buf, buf + k, 0, (size_t)c, 256 * 1024};
        worker(&w);
        for (int i = 0; i < k + m && j->rc == 0; i++) {
            int64_t put = 0;
            while (put < c) {
                ssize_t n = pwrite(j->outs[i], buf[i] + put, (size_t)(c - put), (off_t)(off + put));
                if (n < 0) { if (errno == EINTR) continue; j->rc = -errno; break; }
                put += n;
            }
        }
    }
    for (int i = 0; i < k + m; i++) free(buf[i]);
    return NULL;
}

int orc_generate_ec_files_mt(const char *base, int kind, int threads, int64_t large, int64_t small, int k, int m)
{
    if (kind == 0 && !g_ref_handle) return -ENOSYS;
    if (kind == 1 && !orc_cpu_has_gfni()) return -ENOSYS;
    if (k <= 0 || m <= 0 || k + m > ORC_MAX_SHARDS || large <= 0 || small <= 0 || large % small || small % 4096) return -EINVAL;
    if (threads < 1) threads = 1;
    char path[4096];
    snprintf(path, sizeof path, "%s.dat", base);
    const int dat = open(path, O_RDONLY);
    if (dat < 0) return -errno;
    struct stat st;
    fstat(dat, &st);
    int outs[ORC_MAX_SHARDS], rc = 0;
    for (int i = 0; i < k + m; i++) {
        snprintf(path, sizeof path, "%s.ec%02d", base, i);
        outs[i] = open(path, O_TRUNC | O_CREAT | O_WRONLY, 0644);
        if (outs[i] < 0) rc = -errno;
    }
    uint8_t *gen = (uint8_t *)malloc((size_t)(k + m) * k);
    orc_build_matrix(k, k + m, gen);
    const uint8_t *rows = gen + (size_t)k * k;
    uint64_t *mats = NULL;
    if (kind == 1) {
        mats = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)m * k);
        for (int i = 0; i < m * k; i++) mats[i] = gfni_matrix(rows[i]);
    }
    int64_t chunk = small;
    while (chunk > (1 << 20) && chunk % 2 == 0 && (chunk / 2) % 4096 == 0) chunk /= 2;
    gm_job_t *jobs = (gm_job_t *)calloc((size_t)threads, sizeof(gm_job_t));
    pthread_t *tids = (pthread_t *)calloc((size_t)threads, sizeof(pthread_t));
    if (rc == 0) {
        for (int t = 0; t < threads; t++) {
            jobs[t] = (gm_job_t){.kind = kind, .k = k, .m = m, .id = t, .threads = threads, .dat = dat, .outs = outs,
                                 .dat_size = st.st_size, .large = large, .small = small,
                                 .shard_size = orc_expected_shard_size(st.st_size, k, large, small), .chunk = chunk,
                                 .rows = rows, .mats = mats};
            pthread_create(&tids[t], NULL, gm_worker, &jobs[t]);
        }
        for (int t = 0; t < threads; t++) {
            pthread_join(tids[t], NULL);
            if (jobs[t].rc && !rc) rc = jobs[t].rc;
        }
    }
    for (int i = 0; i < k + m; i++)
        if (outs[i] >= 0) close(outs[i]);
    close(dat);
    free(jobs);
    free(tids);
    free(mats);
    free(gen);
    return rc;
}
