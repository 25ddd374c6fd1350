/* scripts/experiments/io_probe.c — MEASUREMENT ONLY.  The I/O schedule of swec_generate_ec_files without the GPU:
 * for every 8 MiB-per-shard stripe, 10 preads of the .dat (one per data block) then 14 pwrites (10 data shards
 * verbatim + 4 "parity" shards — here a copy of block 0-3, the arithmetic is not the point), stripes pipelined
 * DEPTH deep, every batch of preads / pwrites spread over a pool of T threads.  Answers: what can this box's
 * page cache / tmpfs / disk do for this access pattern, i.e. how far is the 2.5-4.2 GB/s of the real pipeline
 * (profiles/r01u_*, r01r_*) from the ceiling, and does O_DIRECT or a deeper pipeline move it?
 *   io_probe DIR GIB THREADS DEPTH DIRECT(0|1) PREALLOC(0|1)                                                  */
#define _GNU_SOURCE
#include <errno.h>
#include <fcntl.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#define K 10
#define M 4
#define CHUNK ((size_t)8 << 20)

typedef struct { int fd; int write; uint8_t *buf; size_t len; off_t off; } Op;
static Op *ops; static int n_ops, next_op, done_ops, stop;
static pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
static pthread_cond_t cv = PTHREAD_COND_INITIALIZER, cv_done = PTHREAD_COND_INITIALIZER;

static void run_op(const Op *o) {
    size_t got = 0;
    while (got < o->len) {
        ssize_t n = o->write ? pwrite(o->fd, o->buf + got, o->len - got, o->off + (off_t)got)
                             : pread(o->fd, o->buf + got, o->len - got, o->off + (off_t)got);
        if (n < 0) { if (errno == EINTR) continue; perror(o->write ? "pwrite" : "pread"); exit(1); }
        if (n == 0) { memset(o->buf + got, 0, o->len - got); break; }
        got += (size_t)n;
    }
}
static void *worker(void *arg) {
    (void)arg;
    pthread_mutex_lock(&mu);
    for (;;) {
        while (!stop && next_op >= n_ops) pthread_cond_wait(&cv, &mu);
        if (stop) break;
        Op o = ops[next_op++];
        pthread_mutex_unlock(&mu);
        run_op(&o);
        pthread_mutex_lock(&mu);
        if (++done_ops == n_ops) pthread_cond_broadcast(&cv_done);
    }
    pthread_mutex_unlock(&mu);
    return NULL;
}
static void submit_and_wait(Op *batch, int n) {   /* one batch at a time per caller; callers are serialised by design */
    pthread_mutex_lock(&mu);
    ops = batch; n_ops = n; next_op = 0; done_ops = 0;
    pthread_cond_broadcast(&cv);
    while (done_ops < n) pthread_cond_wait(&cv_done, &mu);
    n_ops = 0;
    pthread_mutex_unlock(&mu);
}
static double now(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + t.tv_nsec * 1e-9; }

int main(int argc, char **argv) {
    if (argc < 7) { fprintf(stderr, "usage: %s DIR GIB THREADS DEPTH DIRECT PREALLOC\n", argv[0]); return 2; }
    const char *dir = argv[1];
    const size_t gib = (size_t)atoll(argv[2]);
    const int threads = atoi(argv[3]), depth = atoi(argv[4]), direct = atoi(argv[5]), prealloc = atoi(argv[6]);
    const size_t block = ((gib << 30) / K) & ~(CHUNK - 1), dat_size = block * K;   /* one row of K blocks, O_DIRECT-aligned */
    char path[512];
    snprintf(path, sizeof path, "%s/ioprobe.dat", dir);
    int dfd = open(path, O_CREAT | O_TRUNC | O_RDWR, 0644);
    if (dfd < 0) { perror("open dat"); return 1; }
    uint8_t *fill = NULL;
    if (posix_memalign((void **)&fill, 4096, CHUNK)) return 1;
    for (size_t i = 0; i < CHUNK; i += 8) *(uint64_t *)(fill + i) = i * 0x9E3779B97F4A7C15ull;
    for (size_t off = 0; off < dat_size; off += CHUNK) if (pwrite(dfd, fill, CHUNK, (off_t)off) != (ssize_t)CHUNK) { perror("fill"); return 1; }
    close(dfd);
    const int flags = direct ? O_DIRECT : 0;
    dfd = open(path, O_RDONLY | flags);
    if (dfd < 0) { perror("open dat (direct?)"); return 1; }
    int out[K + M];
    for (int i = 0; i < K + M; i++) {
        snprintf(path, sizeof path, "%s/ioprobe.ec%02d", dir, i);
        out[i] = open(path, O_CREAT | O_TRUNC | O_WRONLY | flags, 0644);
        if (out[i] < 0) { perror("open shard"); return 1; }
        if (prealloc && posix_fallocate(out[i], 0, (off_t)block) != 0) { /* best effort */ }
    }
    pthread_t th[64];
    for (int t = 0; t < threads && t < 64; t++) pthread_create(&th[t], NULL, worker, NULL);
    uint8_t **slot = calloc((size_t)depth, sizeof *slot);
    for (int d = 0; d < depth; d++) if (posix_memalign((void **)&slot[d], 4096, (K + M) * CHUNK)) return 1;

    /* DEPTH stripes in flight: software pipeline — batch = reads of stripe s + writes of stripe s-1 (the pool
     * sees 24 independent ops at once, as the real pipeline's reader and writer threads do together) */
    const size_t stripes = block / CHUNK;
    Op *batch = calloc((size_t)(K + M) * (size_t)depth * 2, sizeof *batch);
    const double t0 = now();
    for (size_t s = 0; s < stripes + (size_t)depth - 1; s += (size_t)(depth > 1 ? depth - 1 : 1)) {
        int n = 0;
        for (int d = 0; d < (depth > 1 ? depth - 1 : 1); d++) {
            const size_t rs = s + (size_t)d;                               /* stripe to read */
            if (rs < stripes) {
                uint8_t *b = slot[rs % (size_t)depth];
                for (int i = 0; i < K; i++) batch[n++] = (Op){dfd, 0, b + (size_t)i * CHUNK, CHUNK, (off_t)((size_t)i * block + rs * CHUNK)};
            }
            if (rs >= 1 && rs - 1 < stripes && depth > 1) {                /* stripe to write: read in the previous round */
                const size_t ws = rs - 1;
                uint8_t *b = slot[ws % (size_t)depth];
                for (int i = 0; i < K + M; i++)
                    batch[n++] = (Op){out[i], 1, b + (size_t)(i < K ? i : i - K) * CHUNK, CHUNK, (off_t)(ws * CHUNK)};
            }
        }
        if (depth == 1 && s < stripes) {                                   /* strictly serial: read, then write */
            submit_and_wait(batch, n);
            n = 0;
            for (int i = 0; i < K + M; i++) batch[n++] = (Op){out[i], 1, slot[0] + (size_t)(i < K ? i : i - K) * CHUNK, CHUNK, (off_t)(s * CHUNK)};
        }
        if (n) submit_and_wait(batch, n);
    }
    const double dt = now() - t0;
    pthread_mutex_lock(&mu); stop = 1; pthread_cond_broadcast(&cv); pthread_mutex_unlock(&mu);
    for (int t = 0; t < threads && t < 64; t++) pthread_join(th[t], NULL);
    printf("{\"dir\": \"%s\", \"dat_GiB\": %zu, \"threads\": %d, \"depth\": %d, \"o_direct\": %d, \"prealloc\": %d, "
           "\"seconds\": %.3f, \"dat_GBps\": %.2f, \"total_io_GBps\": %.2f}\n",
           dir, gib, threads, depth, direct, prealloc, dt, dat_size / dt / 1e9, 2.4 * dat_size / dt / 1e9);
    for (int i = 0; i < K + M; i++) { close(out[i]); snprintf(path, sizeof path, "%s/ioprobe.ec%02d", dir, i); unlink(path); }
    snprintf(path, sizeof path, "%s/ioprobe.dat", dir); unlink(path);
    return 0;
}
