/* tests/c/cgo_shaped_harness.c — a plain-C11 caller of include/swec.h that makes the calls the cgo shim
 * of INTEGRATION.md makes, in the same shapes (Go is not installed here, so C stands in for cgo's C side):
 *
 *   abi                       no GPU needed: every host-only entry point the shim touches before its
 *                             first compute call, and the error mapping of a compute call without a device
 *   encode IN OUT N           Encoder.Encode as encodeDataOneBatch issues it (ec_encoder.go:248-278): 14
 *                             malloc'ed (pageable, Go-heap-like) buffers of N bytes, data read from IN
 *                             (10*N bytes), parity appended to OUT; then Encoder.Reconstruct with four nil
 *                             shards (rebuildEcFiles, ec_encoder.go:340-376) and ReconstructData with one
 *                             (store_ec.go:551), each compared with the originals; then the same Encode on the
 *                             pinned batch buffers of swecAllocShardBuffers (ONE swec_alloc_pinned_for_device
 *                             allocation cut into 14 slices: zero-copy kernel / strided DMA), the device order
 *                             swecPickDevice walks, and the decode-kernel cache statistics
 *
 * Compiled by tests/test_c_harness.py with  gcc -std=c11 -Wall -Wextra -Werror -pedantic  — which is also
 * the check that the header is valid C, not just C++.  Test infrastructure; not part of the product.   */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "swec.h"

#define K 10
#define M 4
#define T (K + M)

static int fails;
#define CHECK(cond)                                                          \
    do {                                                                     \
        if (!(cond)) {                                                       \
            fprintf(stderr, "%s:%d: CHECK(%s) failed — last error: %s\n", __FILE__, __LINE__, #cond, \
                    swec_last_error());                                      \
            fails++;                                                         \
        }                                                                    \
    } while (0)

static int run_abi(void) {
    static const uint8_t want_parity[M][K] = {/* Vandermonde-derived RS(10,4) parity rows */
        {129, 150, 175, 184, 210, 196, 254, 232, 3, 2}, {150, 129, 184, 175, 196, 210, 232, 254, 2, 3},
        {191, 214, 98, 10, 6, 111, 223, 183, 5, 4},     {214, 191, 10, 98, 111, 6, 183, 223, 4, 5}};
    CHECK(strstr(swec_version(), "swec") != NULL);
    for (int s = 0; s >= SWEC_ERR_DELETED; s--) CHECK(strcmp(swec_strerror(s), "unknown error") != 0);
    CHECK(strcmp(swec_strerror(-100), "unknown error") == 0);

    swec_encoder *enc = NULL;
    CHECK(swec_encoder_new(0, 4, -1, &enc) == SWEC_ERR_INVALID_ARG && enc == NULL); /* ErrInvShardNum */
    CHECK(swec_encoder_new(30, 3, -1, &enc) == SWEC_ERR_INVALID_ARG);               /* > MaxShardCount */
    CHECK(swec_encoder_new(K, M, -1, &enc) == SWEC_OK && enc != NULL);
    uint8_t gen[T * K];
    CHECK(swec_encoder_matrix(enc, gen) == SWEC_OK);
    for (int r = 0; r < K; r++)
        for (int c = 0; c < K; c++) CHECK(gen[r * K + c] == (r == c));
    CHECK(memcmp(gen + K * K, want_parity, sizeof want_parity) == 0);

    /* shards 0-3 lost: the decode rows of SURVEY §8c */
    static const uint8_t want_row0[K] = {29, 239, 227, 16, 49, 195, 195, 48, 13, 12};
    uint8_t present[T] = {0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1}, rows[T * K];
    int inputs[K], outputs[T], nout = 0;
    CHECK(swec_reconstruct_matrix(enc, present, 0, inputs, outputs, &nout, rows) == SWEC_OK);
    CHECK(nout == 4 && outputs[0] == 0 && outputs[3] == 3 && inputs[0] == 4 && inputs[9] == 13);
    CHECK(memcmp(rows, want_row0, K) == 0);
    memset(present, 0, sizeof present);
    present[0] = 1;
    CHECK(swec_reconstruct_matrix(enc, present, 0, inputs, outputs, &nout, rows) == SWEC_ERR_TOO_FEW_SHARDS);

    /* a compute call on a host-only encoder: the loud no-fallback error the shim turns into a Go error */
    uint8_t *bufs[T];
    for (int i = 0; i < T; i++) bufs[i] = calloc(1, 4096);
    CHECK(swec_encode(enc, bufs, 4096) == SWEC_ERR_NO_DEVICE);
    CHECK(strlen(swec_last_error()) > 0);
    CHECK(swec_encode(enc, bufs, 0) == SWEC_ERR_INVALID_ARG); /* ErrShardNoData */
    for (int i = 0; i < T; i++) free(bufs[i]);
    swec_encoder_free(enc);
    swec_encoder_free(NULL);

    /* layout arithmetic the Go side keeps using (ec_locate.go, disk_location_ec.go:428-448) */
    const int64_t G = (int64_t)1 << 30, Mi = (int64_t)1 << 20;
    CHECK(swec_expected_shard_size(30 * G, K, G, Mi) == 3 * G);                 /* exact multiple: 3 large rows */
    CHECK(swec_expected_shard_size(30 * G + 1, K, G, Mi) == 3 * G + Mi);
    CHECK(swec_expected_shard_size(2590912, K, G, Mi) == Mi);                   /* the fixture volume 1.dat */
    CHECK(swec_expected_shard_size(0, K, G, Mi) == 0);
    swec_interval iv[8];
    int n = swec_locate_data(G, Mi, 3 * G, 10 * G - 5, 10, K, iv, 8);           /* straddles large rows 0/1 */
    CHECK(n == 2 && iv[0].is_large_block && iv[0].block_index == 9 && iv[0].size == 5 && iv[1].block_index == 10);
    int sid = -1;
    int64_t soff = -1;
    swec_interval_to_shard(&iv[1], G, Mi, K, &sid, &soff);
    CHECK(sid == 0 && soff == G);
    CHECK(swec_locate_data(G, Mi, 3 * G, 0, 100, K, iv, 0) == SWEC_ERR_INVALID_ARG);

    /* file-level argument checks happen before any device work */
    CHECK(swec_generate_ec_files("/nonexistent/1", 256 * 1024, G, Mi, K, M, 0) == SWEC_ERR_IO);
    CHECK(swec_generate_ec_files("/nonexistent/1", 3, G, Mi, K, M, 0) == SWEC_ERR_INVALID_ARG);
    uint32_t rebuilt[SWEC_MAX_SHARDS];
    int nre = -1;
    CHECK(swec_rebuild_ec_files("/nonexistent/1", NULL, 0, K, M, 0, rebuilt, &nre) == SWEC_ERR_TOO_FEW_SHARDS);
    CHECK(nre == 0);
    /* round-2 entry points, host-only behaviour */
    int aot = 0;
    CHECK(swec_jit_stats(NULL, NULL, &aot, NULL) == SWEC_OK && aot == 15);
    CHECK(swec_set_option("host_zero_copy", 2) == SWEC_OK && swec_set_option("host_zero_copy", 3) == SWEC_ERR_INVALID_ARG);
    CHECK(swec_set_option("file_direct_io", 0) == SWEC_OK && swec_set_option("power_mode", 0) == SWEC_OK);
    swec_shutdown();
    swec_shutdown(); /* idempotent */
    return fails;
}

static uint8_t *slurp(const char *path, size_t want) {
    FILE *f = fopen(path, "rb");
    if (!f) return NULL;
    uint8_t *p = malloc(want);
    if (p && fread(p, 1, want, f) != want) {
        free(p);
        p = NULL;
    }
    fclose(f);
    return p;
}

static int run_encode(const char *in_path, const char *out_path, size_t n) {
    uint8_t *data = slurp(in_path, K * n);
    if (!data) {
        fprintf(stderr, "cannot read %zu bytes from %s\n", K * n, in_path);
        return 2;
    }
    int ndev = 0;
    CHECK(swec_device_count(&ndev) == SWEC_OK && ndev >= 1);
    swec_encoder *enc = NULL;
    CHECK(swec_encoder_new(K, M, 0, &enc) == SWEC_OK);
    if (fails) return fails;

    uint8_t *shards[T], *orig[T];
    for (int i = 0; i < T; i++) {
        shards[i] = malloc(n);
        orig[i] = malloc(n);
        if (i < K) memcpy(shards[i], data + (size_t)i * n, n);
        else memset(shards[i], 0xEE, n); /* parity slices are overwritten, whatever they held */
    }
    CHECK(swec_encode(enc, shards, n) == SWEC_OK);
    for (int i = 0; i < K; i++) CHECK(memcmp(shards[i], data + (size_t)i * n, n) == 0); /* data untouched */
    int ok = 0;
    CHECK(swec_verify(enc, shards, n, &ok) == SWEC_OK && ok == 1);
    FILE *out = fopen(out_path, "wb");
    CHECK(out != NULL);
    for (int i = K; i < T && out; i++) CHECK(fwrite(shards[i], 1, n, out) == n);
    if (out) fclose(out);
    for (int i = 0; i < T; i++) memcpy(orig[i], shards[i], n);

    /* Reconstruct: nil shards 0, 3, 11, 13 — the shim allocates them, as klauspost does */
    uint8_t present[T];
    memset(present, 1, sizeof present);
    const int lost[4] = {0, 3, 11, 13};
    for (int j = 0; j < 4; j++) {
        present[lost[j]] = 0;
        memset(shards[lost[j]], 0x55, n);
    }
    CHECK(swec_reconstruct(enc, shards, present, n, 0) == SWEC_OK);
    for (int i = 0; i < T; i++) CHECK(memcmp(shards[i], orig[i], n) == 0);

    /* ReconstructData: a parity shard and a data shard missing, only the data shard is filled */
    memset(present, 1, sizeof present);
    present[5] = present[12] = 0;
    memset(shards[5], 0x11, n);
    memset(shards[12], 0x22, n);
    CHECK(swec_reconstruct(enc, shards, present, n, 1) == SWEC_OK);
    CHECK(memcmp(shards[5], orig[5], n) == 0);
    CHECK(shards[12][0] == 0x22 && shards[12][n - 1] == 0x22);

    /* a corrupted byte is caught by Verify */
    shards[12][0] = orig[12][0];
    memcpy(shards[12], orig[12], n);
    shards[2][n / 2] ^= 0x40;
    CHECK(swec_verify(enc, shards, n, &ok) == SWEC_OK && ok == 0);

    /* five lost: ErrTooFewShards, nothing written */
    memset(present, 1, sizeof present);
    for (int i = 0; i < 5; i++) present[i] = 0;
    CHECK(swec_reconstruct(enc, shards, present, n, 0) == SWEC_ERR_TOO_FEW_SHARDS);

    /* swecAllocShardBuffers (integration/go/ec_swec.go): the batch buffers of encodeDatFile / rebuildEcFiles as 14
     * slices of ONE pinned allocation at a 4 KiB-rounded pitch; Encode on them must give the pageable call's parity,
     * and a single lost shard must come back through the compiled-in reconstruct kernel (no NVRTC) */
    {
        const size_t pitch = (n + 4095) & ~(size_t)4095;
        uint8_t *base = swec_alloc_pinned_for_device(0, T * pitch);
        CHECK(base != NULL);
        if (base) {
            uint8_t *pin[T];
            for (int i = 0; i < T; i++) {
                pin[i] = base + (size_t)i * pitch;
                if (i < K) memcpy(pin[i], data + (size_t)i * n, n);
                else memset(pin[i], 0xEE, n);
            }
            CHECK(swec_encode(enc, pin, n) == SWEC_OK);
            for (int i = 0; i < T; i++) CHECK(memcmp(pin[i], orig[i], n) == 0);
            uint64_t aot0 = 0, aot1 = 0;
            int aot_matrices = 0;
            CHECK(swec_jit_stats(NULL, NULL, &aot_matrices, &aot0) == SWEC_OK && aot_matrices == 15);
            memset(present, 1, sizeof present);
            present[7] = 0;
            memset(pin[7], 0x33, n);
            CHECK(swec_reconstruct(enc, pin, present, n, 1) == SWEC_OK);
            CHECK(memcmp(pin[7], orig[7], n) == 0);
            CHECK(swec_jit_stats(NULL, NULL, NULL, &aot1) == SWEC_OK);
            CHECK(aot1 == aot0 + 1); /* one launch of the compiled-in single-loss kernel (a <16 B tail adds the byte kernel) */
            swec_free_pinned(base);
        }
        int order[64], cnt = 0;
        CHECK(swec_device_spread_order(order, 64, &cnt) == SWEC_OK && cnt == ndev);
        int seen = 0;
        for (int i = 0; i < cnt; i++) seen += order[i] >= 0 && order[i] < ndev;
        CHECK(seen == ndev);
        double heat = -1;
        int lp = -1;
        CHECK(swec_debug_power_state(0, &heat, &lp) == SWEC_OK && heat >= 0 && (lp == 0 || lp == 1));
    }

    CHECK(swec_kernel_launches() > 0);
    swec_encoder_free(enc);
    for (int i = 0; i < T; i++) {
        free(shards[i]);
        free(orig[i]);
    }
    free(data);
    return fails;
}

int main(int argc, char **argv) {
    if (argc == 2 && strcmp(argv[1], "abi") == 0) return run_abi() ? 1 : 0;
    if (argc == 5 && strcmp(argv[1], "encode") == 0) return run_encode(argv[2], argv[3], (size_t)atoll(argv[4])) ? 1 : 0;
    fprintf(stderr, "usage: %s abi | encode IN OUT SHARD_LEN\n", argv[0]);
    return 2;
}
