/*
 * swec.h — C ABI of the B200-native Reed–Solomon erasure-coding engine for SeaweedFS.
 *
 * This is the drop-in boundary for the RS(10,4) hot path of weed/storage/erasure_coding:
 * a thin cgo file binds these entry points in place of
 * github.com/klauspost/reedsolomon (go.mod:50).  Plain pointers and sizes only; every function
 * returns 0 (SWEC_OK) or a negative swec_status, never aborts, never falls back to the CPU: if
 * no CUDA device / kernel image is usable the call fails with SWEC_ERR_NO_DEVICE / SWEC_ERR_CUDA
 * so the caller's cleanup-on-error path runs (weed/server/volume_grpc_erasure_coding.go:78-87).
 *
 * Reference interface each group replaces (paths relative to the SeaweedFS tree):
 *   swec_encoder_new            reedsolomon.New(ds, ps)   weed/storage/erasure_coding/ec_context.go:34-36,
 *                                                          weed/storage/store_ec.go:485
 *   swec_encode                 Encoder.Encode            weed/storage/erasure_coding/ec_encoder.go:265
 *   swec_reconstruct            Encoder.Reconstruct       weed/storage/erasure_coding/ec_encoder.go:360
 *                               Encoder.ReconstructData   weed/storage/store_ec.go:551   (data_only = 1)
 *   swec_verify                 (Rust twin) rs.verify     seaweed-volume/src/storage/erasure_coding/ec_encoder.rs:177-278
 *   swec_write_ec_files         WriteEcFiles / generateEcFiles   ec_encoder.go:61-69,110-128
 *   swec_rebuild_ec_files       RebuildEcFiles / generateMissingEcFiles   ec_encoder.go:74-104,146-200
 *   swec_verify_ec_files        (Rust twin) verify_ec_shards   seaweed-volume/src/storage/erasure_coding/ec_encoder.rs:177-278
 *   swec_reconstruct_batch      batched ReconstructData   weed/storage/store_ec.go:482-560 (one call per interval today)
 *   swec_write_dat_file         WriteDatFile              weed/storage/erasure_coding/ec_decoder.go:176-223
 *   swec_ec_shards_generate     VolumeEcShardsGenerate (file work)   weed/server/volume_grpc_erasure_coding.go:43-146
 *   swec_ec_shards_rebuild      VolumeEcShardsRebuild  (file work)   weed/server/volume_grpc_erasure_coding.go:149-225
 *   swec_ec_shards_to_volume    VolumeEcShardsToVolume (file work)   weed/server/volume_grpc_erasure_coding.go:578-668
 *   swec_read_ec_needles        Store.ReadEcShardNeedle (local shards, batched)   weed/storage/store_ec.go:252-355,482-560
 *   swec_check_index_file       idx.CheckIndexFile / EcVolume.ScrubIndex   weed/storage/idx/check.go:36-111
 *   swec_ec_volume_*            EcVolume: mount, ReadEcShardNeedle, DeleteNeedleFromEcx, FileAndDeleteCount, ScrubLocal
 *                                                          weed/storage/erasure_coding/ec_volume.go, ec_volume_delete.go, ec_volume_scrub.go
 *   swec_locate_data            LocateData                weed/storage/erasure_coding/ec_locate.go:16-53
 *   swec_expected_shard_size    calculateExpectedShardSize   weed/storage/disk_location_ec.go:428-448
 *
 * Threading: every entry point may be called concurrently from any OS thread (Go schedules
 * gRPC handlers freely; the shell runs up to 10 volumes at once, weed/shell/common.go:11).
 * Calls on ONE encoder handle serialise on its staging buffers; use one handle per goroutine
 * (as the reference does: one reedsolomon.Encoder per file) for parallelism.
 * Ownership: the caller owns every buffer; nothing is retained after a call returns.
 */
#ifndef SWEC_H
#define SWEC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SWEC_MAX_SHARDS 32 /* MaxShardCount, ec_encoder.go:23 */

typedef enum swec_status {
    SWEC_OK = 0,
    SWEC_ERR_INVALID_ARG = -1,     /* bad shard counts, NULL pointers, mismatched sizes        */
    SWEC_ERR_TOO_FEW_SHARDS = -2,  /* fewer than data_shards present (ErrTooFewShards)          */
    SWEC_ERR_CUDA = -3,            /* a CUDA call failed; swec_last_error() has the detail      */
    SWEC_ERR_IO = -4,              /* open/read/write failed; errno text in swec_last_error()   */
    SWEC_ERR_NOMEM = -5,
    SWEC_ERR_SHARD_SIZE = -6,      /* shard files of unequal length ("ec shard size expected…") */
    SWEC_ERR_NO_DEVICE = -7,       /* no usable CUDA device — there is NO CPU fallback          */
    SWEC_ERR_JIT = -8,             /* run-time kernel specialisation failed                     */
    SWEC_ERR_NO_LIVE_NEEDLES = -9, /* ec.decode of a volume whose index has no live entries      */
    SWEC_ERR_NOT_FOUND = -10,      /* needle id not in .ecx (erasure_coding.NotFoundError)       */
    SWEC_ERR_DELETED = -11         /* needle is tombstoned or journalled (storage.ErrorDeleted)  */
} swec_status;

typedef struct swec_encoder swec_encoder;

/* ---- library ------------------------------------------------------------------------------ */
const char *swec_version(void);
const char *swec_strerror(int status);
const char *swec_last_error(void);            /* thread-local detail of the last failure        */
int swec_device_count(int *count);            /* SWEC_ERR_NO_DEVICE when the driver is absent   */
/* Stop the library's background compiler thread and release the staging rings that file-level calls park
 * for the next call (idempotent).  Embedders whose runtime tears the
 * process down in stages (CPython's Py_Finalize) call this from their own exit hook; the library
 * also registers it with atexit().  Everything keeps working afterwards, without background JIT. */
void swec_shutdown(void);
/* Device ids interleaved over the host's NUMA nodes (0,4,1,5,… on a 2-socket, 8-GPU box): the first n entries are the
 * n GPUs a process should use for n concurrent host-fed volumes — one socket cannot feed four GPUs at full PCIe rate.
 * This is the order swecPickDevice walks; weed/shell/command_ec_encode.go:302-315 runs the volumes.   */
int swec_device_spread_order(int *order, int capacity, int *count);
uint64_t swec_kernel_launches(void);          /* kernels this process has launched (all devices) */
/* Tuning: "enc_threads" {128,256,512}, "enc_unroll" {1,2}, "ctas_per_sm" (0 = auto),
 * "stage_chunk" (bytes per shard per staging slot), "stage_slots", "host_pieces" (a host-buffer call is cut into at
 * least this many pipelined pieces), "host_min_chunk" (but none smaller than this many bytes per shard), "jit" {0,1},
 * "jit_min_bytes" (streams at least this long compile their kernel inline, shorter ones in the
 * background), "power_mode" {0 = auto by the device's recent kernel time (default) — a GPU that runs
 * encode launches back to back for more than ~0.4 s sits on its power cap, where the variant with fewer instructions is
 * 4-5 % faster; 1 = always the boost-clock kernel variant, 2 = always the low-power one}, "host_zero_copy" {0,1,2 = auto:
 * host-buffer calls of at most "host_zero_copy_max" bytes per shard run the kernel directly on the (mapped, pinned) host
 * memory over PCIe instead of staging through HBM}, "file_direct_io" {bit 0: O_DIRECT reads, bit 1: O_DIRECT writes in
 * the file-level entry points}.  Measurement knobs: "xt_variant" {0..3} (instruction mix of run-time specialised kernels,
 * device_common.cuh), "use_aot" {0,1} (0: RS(10,4) encode is specialised at run time like any matrix), "jit_share_powers"
 * {0,1} (1: run-time specialised kernels are generated with shared power chains — fewer multiply-by-2 steps; CPU-verified,
 * not yet measured on a B200, hence off).                                                                          */
int swec_set_option(const char *name, long value);
/* Diagnostics: generate and NVRTC-compile (sm_100a) the specialised kernel for an r×k matrix without
 * loading it — needs no GPU.  Reports the cubin size and the generator's instruction statistics.  */
int swec_debug_jit_compile(int r, int k, const uint8_t *rows, size_t *cubin_bytes, int *xtime_steps,
                           int *xor_ops);
/* Diagnostics of "power_mode" auto: `heat_ms` = estimated Horner-kernel milliseconds the device ran during the last
 * second (exponentially decayed; continuous encoding tends to 1000), `low_power` = the variant the next launch takes
 * (1 once heat_ms exceeds 450).  device < 0: the calling thread's current device.                                  */
int swec_debug_power_state(int device, double *heat_ms, int *low_power);
/* The decode-kernel cache (GPU analogue of the decode-matrix LRU, rse/src/core.rs:25,700-734), three tiers:
 * `aot_matrices` reconstruct matrices compiled with the library (every single-shard loss of RS(10,4) and shards 0-3
 * lost: no compile, no NVRTC, any stream length); an on-disk cubin cache shared by every process
 * ($SWEC_CACHE_DIR, else $XDG_CACHE_HOME/swec, else ~/.cache/swec; SWEC_NO_DISK_CACHE=1 disables; a directory that does
 * not belong to the calling user or that group/others can write to is not trusted with executable code and is ignored) whose hits are
 * counted in `disk_cache_hits`; NVRTC for patterns never seen before (`nvrtc_compiles`).  `aot_launches` = launches of
 * the compiled-in reconstruct kernels by this process.  Any pointer may be NULL. */
int swec_jit_stats(uint64_t *nvrtc_compiles, uint64_t *disk_cache_hits, int *aot_matrices,
                   uint64_t *aot_launches);

/* ---- encoder = reedsolomon.New(dataShards, parityShards) ----------------------------------- */
/* device < 0: host-side object only (matrix queries); compute calls then fail with NO_DEVICE.  */
int swec_encoder_new(int data_shards, int parity_shards, int device, swec_encoder **out);
void swec_encoder_free(swec_encoder *enc);
/* (k+m)×k generator matrix, row-major: identity on top, parity rows below.                     */
int swec_encoder_matrix(const swec_encoder *enc, uint8_t *out);
/* The fused matrix Reconstruct would apply for a presence mask: inputs[k] (shard ids read),
 * outputs[*n_outputs] (shard ids produced), rows[*n_outputs * k].                              */
int swec_reconstruct_matrix(const swec_encoder *enc, const uint8_t *present, int data_only,
                            int *inputs, int *outputs, int *n_outputs, uint8_t *rows);

/* ---- Encoder.Encode / Reconstruct / ReconstructData on HOST buffers ------------------------- */
/* shards[0..k) are read, shards[k..k+m) are overwritten; all shard_len bytes, shard_len > 0.   */
int swec_encode(swec_encoder *enc, uint8_t *const *shards, size_t shard_len);
/* present[i] != 0 ⇔ shards[i] holds data.  Missing shards must point at shard_len writable
 * bytes (the cgo shim allocates them, as klauspost does for nil slices).  data_only = 1 fills
 * only indices < k (ReconstructData).  All present ⇒ no-op; fewer than k ⇒ TOO_FEW_SHARDS.     */
int swec_reconstruct(swec_encoder *enc, uint8_t *const *shards, const uint8_t *present,
                     size_t shard_len, int data_only);
/* Many ReconstructData/Reconstruct calls at once — the batched form of the degraded-read path
 * (store_ec.go:482-560 issues one tiny call per needle interval).  Items that share a presence
 * mask are packed into common launches; results are identical to calling swec_reconstruct on each. */
typedef struct swec_reconstruct_item {
    uint8_t *const *shards;   /* k+m pointers, as for swec_reconstruct */
    const uint8_t *present;   /* k+m flags */
    size_t shard_len;
    int data_only;
} swec_reconstruct_item;
int swec_reconstruct_batch(swec_encoder *enc, const swec_reconstruct_item *items, int n_items);
/* One call, several GPUs: the byte-column range [0, shard_len) is cut into n_encs contiguous pieces and
 * piece g goes through encs[g] (one handle per GPU, created with swec_encoder_new(k, m, device_g)), all
 * pieces concurrently, each over its own GPU's PCIe link.  Columns are independent, so the result is
 * byte-identical to swec_encode / swec_reconstruct on one handle; there is no inter-GPU traffic.  This
 * is the second axis of independence of SURVEY §8e (the first — whole volumes round-robin over GPUs —
 * needs no API: give each concurrent volume a handle on another device).                            */
int swec_encode_multi(swec_encoder *const *encs, int n_encs, uint8_t *const *shards, size_t shard_len);
int swec_reconstruct_multi(swec_encoder *const *encs, int n_encs, uint8_t *const *shards,
                           const uint8_t *present, size_t shard_len, int data_only);
/* n_shards pinned buffers of shard_len bytes laid out FOR such a call: byte range g of every shard is bound to the
 * NUMA node of encs[g]'s GPU (the split rule is shared), so each GPU DMAs from its own socket.  One allocation:
 * release it with swec_free_pinned(shards[0]).                                                                */
int swec_alloc_pinned_shards(swec_encoder *const *encs, int n_encs, int n_shards, size_t shard_len,
                             uint8_t **shards);
/* *ok = 1 iff the parity shards match the data shards.                                         */
int swec_verify(swec_encoder *enc, uint8_t *const *shards, size_t shard_len, int *ok);

/* ---- the same on DEVICE buffers, asynchronous on `stream` (a cudaStream_t; NULL = the CUDA
 *      default stream).  Buffers must stay valid until the stream reaches this point. --------- */
int swec_encode_device(swec_encoder *enc, const void *const *data, void *const *parity,
                       size_t shard_len, void *stream);
int swec_reconstruct_device(swec_encoder *enc, void *const *shards, const uint8_t *present,
                            size_t shard_len, int data_only, void *stream);
/* The primitive underneath: out[p][x] = XOR_i rows[p*k+i] ⊗ in[i][x] for any r×k matrix over
 * GF(2^8)/0x11D (k ≤ 32) — what code_some_slices does (reed-solomon-erasure core.rs:484-512).    */
int swec_apply_device(swec_encoder *enc, int r, int k, const uint8_t *rows, const void *const *in,
                      void *const *out, size_t shard_len, void *stream);
/* A whole volume image resident in HBM → parity shard images, following the two-tier striping
 * of encodeDatFile (ec_encoder.go:280-321): rows of k large blocks while ≥ k*large bytes
 * remain, then rows of k small blocks, the last one zero-padded.  parity[p] receives
 * swec_expected_shard_size() bytes.  Data shards are views of the image and are not copied.    */
int swec_encode_volume_device(swec_encoder *enc, const void *dat, int64_t dat_size,
                              int64_t large_block, int64_t small_block, void *const *parity,
                              void *stream);
/* Gather data shard `shard_id` of the image into a contiguous shard buffer (what .ecNN holds).  */
int swec_extract_data_shard_device(swec_encoder *enc, const void *dat, int64_t dat_size,
                                   int64_t large_block, int64_t small_block, int shard_id,
                                   void *shard_out, void *stream);
/* WriteDatFile on device memory (weed/storage/erasure_coding/ec_decoder.go:176-223, the body of ec.decode): the k data
 * shards, each swec_expected_shard_size() bytes in HBM (as read from .ec00-.ec09, or as swec_reconstruct_device just
 * rebuilt them), are un-striped into the dat_size bytes of the volume image — rows of k large blocks while at least one
 * full large row remains, then rows of k small blocks, the zero padding of the last row dropped.  Exact inverse of
 * swec_extract_data_shard_device; k strided device-to-device copies per region, asynchronous on `stream`.           */
int swec_write_dat_device(swec_encoder *enc, const void *const *data_shards, int64_t dat_size,
                          int64_t large_block, int64_t small_block, void *dat_out, void *stream);
int swec_stream_synchronize(swec_encoder *enc, void *stream);

/* ---- file level: .dat → .ec00…, missing .ecNN ← the others, .ec00–.ec09 → .dat -------------- */
/* WriteEcFiles(baseFileName): RS(10,4), 1 GiB / 1 MiB blocks.                                   */
int swec_write_ec_files(const char *base_file_name, int device);
/* generateEcFiles(base, bufferSize, large, small, ctx).  buffer_size only has to divide both
 * block sizes (the reference Fatalf's otherwise); results do not depend on it.                  */
int swec_generate_ec_files(const char *base_file_name, int64_t buffer_size, int64_t large_block,
                           int64_t small_block, int data_shards, int parity_shards, int device);
/* RebuildEcFiles(base, additionalDirs...).  data_shards = 0 ⇒ take the ratio from base.vif
 * (ecShardConfig) when valid, else 10+4 (ec_encoder.go:76-95).  rebuilt[] (≥ SWEC_MAX_SHARDS
 * entries) receives the generated shard ids.                                                    */
int swec_rebuild_ec_files(const char *base_file_name, const char *const *additional_dirs,
                          int n_additional_dirs, int data_shards, int parity_shards, int device,
                          uint32_t *rebuilt, int *n_rebuilt);
/* Scrub: re-encode .ec00–.ec(k-1) on the GPU and compare with the parity shards on disk (the Rust
 * twin's verify_ec_shards, seaweed-volume/src/storage/erasure_coding/ec_encoder.rs:177-278; the Go
 * scrub never checks parity, ec_volume_scrub.go:27-118).  mismatched_vectors[p] (m entries, may be
 * NULL) = number of differing 16-byte vectors in parity shard p; *ok = 1 iff all are zero.       */
int swec_verify_ec_files(const char *base_file_name, const char *const *additional_dirs,
                         int n_additional_dirs, int data_shards, int parity_shards, int device,
                         uint64_t *mismatched_vectors, int *ok);
int swec_write_dat_file(const char *base_file_name, int64_t dat_file_size,
                        const char *const *shard_file_names, int data_shards,
                        int64_t large_block, int64_t small_block);

/* ---- whole-volume operations: what the three EC gRPC handlers do to files, in their order --------- */
/* VolumeEcShardsGenerate: ratio from data_base.vif when valid else 10+4; index_base.idx → .ecx FIRST;
 * snapshot the .dat size; .dat → .ec00… on the GPU (256 KiB / 1 GiB / 1 MiB); write data_base.vif
 * {version, datFileSize, expireAtSec, ecShardConfig}.  On any error the shard files and the .ecx are
 * removed (the handler's deferred cleanup).  index_base NULL/"" = data_base.  needle_version 0 = read
 * it from the .dat superblock.                                                                      */
int swec_ec_shards_generate(const char *data_base_file_name, const char *index_base_file_name,
                            uint32_t needle_version, uint64_t expire_at_sec, int device);
/* VolumeEcShardsRebuild: RebuildEcFiles(data_base, additional_dirs...) then RebuildEcxFile(index_base,
 * or data_base when index_base has no .ecx).                                                        */
int swec_ec_shards_rebuild(const char *data_base_file_name, const char *index_base_file_name,
                           const char *const *additional_dirs, int n_additional_dirs, int device,
                           uint32_t *rebuilt, int *n_rebuilt);
/* VolumeEcShardsToVolume (ec.decode): needs all data shards (data_base's directory, then
 * additional_dirs); RebuildEcxFile → HasLiveNeedles (SWEC_ERR_NO_LIVE_NEEDLES when none) →
 * FindDatFileSize → WriteDatFile → WriteIdxFileFromEcIndex.  No GPU work.  *dat_file_size may be NULL. */
int swec_ec_shards_to_volume(const char *data_base_file_name, const char *index_base_file_name,
                             const char *const *additional_dirs, int n_additional_dirs,
                             int64_t *dat_file_size);

/* Store.ReadEcShardNeedle for MANY needles of one EC volume whose shard files are local (data_base's
 * directory, then additional_dirs) — weed/storage/store_ec.go:252-355,482-560: find each needle in .ecx
 * (journalled ids read as deleted), LocateData its record through the two-tier layout (shard size from
 * .vif's datFileSize, else shard file size - 1), pread every interval whose shard file is present and
 * recover the others from the same interval of all remaining shards with ReconstructData.  All
 * recoveries of the call go to the GPU as ONE swec_reconstruct_batch; with every needed shard present the
 * call does no GPU work at all.  buf receives the raw record bytes exactly as the reference's `bytes`
 * (which over-reads: GetActualSize is applied twice, ec_volume.go:395,414); parsing them is the storage
 * engine's business.  Per-needle outcome in status: SWEC_OK, SWEC_ERR_NOT_FOUND, SWEC_ERR_DELETED,
 * SWEC_ERR_TOO_FEW_SHARDS, or SWEC_ERR_INVALID_ARG when capacity < the n_bytes reported back.        */
typedef struct swec_needle_read {
    uint64_t needle_id;            /* in  */
    uint8_t *buf;                  /* in  */
    size_t capacity;               /* in  */
    int64_t offset;                /* out: byte offset of the record in the .dat                  */
    int32_t size;                  /* out: Size field of the index entry (negative = deleted)      */
    int32_t status;                /* out */
    size_t n_bytes;                /* out: bytes stored in buf (or needed, on SWEC_ERR_INVALID_ARG) */
    int32_t n_recovered_intervals; /* out: intervals that were reconstructed rather than read      */
    int32_t reserved;
} swec_needle_read;
int swec_read_ec_needles(const char *data_base_file_name, const char *index_base_file_name,
                         const char *const *additional_dirs, int n_additional_dirs,
                         swec_needle_read *reads, int n_reads, int device);
/* The same on a MOUNTED volume — the twin of the long-lived EcVolume (ec_volume.go:36-160): open once
 * (ratio / needle version / datFileSize from .vif, shard files opened, .ecx loaded, .ecj re-read when it
 * grows), read many times; the encoder behind the recoveries, its staging ring and its specialised kernels
 * live as long as the handle.  Calls on one handle serialise.  swec_read_ec_needles = open + read + close. */
typedef struct swec_ec_volume swec_ec_volume;
int swec_ec_volume_open(const char *data_base_file_name, const char *index_base_file_name,
                        const char *const *additional_dirs, int n_additional_dirs, int device,
                        swec_ec_volume **out);
int swec_ec_volume_read_needles(swec_ec_volume *vol, swec_needle_read *reads, int n_reads);
/* EcVolume.DeleteNeedleFromEcx (ec_volume_delete.go:28-93): append the id to the .ecj journal (fsync'ed
 * before it becomes visible to reads); unknown, tombstoned or already journalled ids are not errors.  */
int swec_ec_volume_delete_needle(swec_ec_volume *vol, uint64_t needle_id);
/* EcVolume.ScrubLocal (ec_volume_scrub.go:27-118) without the needle parse: ScrubIndex (swec_check_index_file on the
 * .ecx), then every live entry is located and each of its chunks read from the local shard that should hold it.
 * broken_shards[SWEC_MAX_SHARDS] receives the ids (ascending) of shards that were too short or unreadable for some
 * chunk; findings are newline-separated in errors[], worded like the reference.  Parity is checked by
 * swec_verify_ec_files, record CRCs by the storage engine.                                                    */
int swec_ec_volume_scrub_local(swec_ec_volume *vol, int64_t *entries, uint32_t *broken_shards, int *n_broken,
                               char *errors, size_t errors_cap, int *n_errors);
/* What mounting derived (NewEcVolume, ec_volume.go:114-154,399-417): EC ratio and needle version from .vif (defaults
 * 10+4, version 3), the shard size LocateData works with, and a bit per shard file found locally.  Any out pointer
 * may be NULL.                                                                                                 */
int swec_ec_volume_info(swec_ec_volume *vol, int *data_shards, int *parity_shards, int *needle_version,
                        int64_t *shard_dat_size, uint32_t *local_shard_bits);
/* EcVolume.FileAndDeleteCount (ec_volume.go:330-349): .ecx entries, distinct journalled deletions.      */
int swec_ec_volume_counts(swec_ec_volume *vol, uint64_t *file_count, uint64_t *delete_count);
void swec_ec_volume_close(swec_ec_volume *vol);

/* ---- index files either side of the path (host only, no GPU) ---------------------------------- */
/* WriteSortedFileFromIdx(base, ext): base.idx → base+ext (".ecx"), live entries sorted by needle id
 * (ec_encoder.go:31-58).  Call it BEFORE writing shards, as VolumeEcShardsGenerate does.          */
int swec_write_sorted_file_from_idx(const char *base_file_name, const char *ext);
/* RebuildEcxFile: fold the .ecj deletion journal into .ecx, then remove .ecj
 * (ec_volume_delete.go:95-142).                                                                  */
int swec_rebuild_ecx_file(const char *base_file_name);
/* WriteIdxFileFromEcIndex: .ecx (+ one tombstone per .ecj id) → .idx (ec_decoder.go:35-60).       */
int swec_write_idx_file_from_ec_index(const char *base_file_name);
/* idx.CheckIndexFile (weed/storage/idx/check.go:36-111) = EcVolume.ScrubIndex on an .ecx (or an .idx): entries
 * processed, and one message per finding — overlapping neighbours in (offset, size) order, a file that is not a whole
 * number of entries — newline-separated into errors[errors_cap] (may be NULL), worded like the reference.        */
int swec_check_index_file(const char *path, int needle_version, int64_t *entries, char *errors,
                          size_t errors_cap, int *n_errors);
/* HasLiveNeedles / FindDatFileSize (ec_decoder.go:23-33, 65-92).                                  */
int swec_has_live_needles(const char *index_base_file_name, int *has_live);
int swec_find_dat_file_size(const char *data_base_file_name, const char *index_base_file_name,
                            int64_t *dat_size);

/* ---- layout arithmetic (no GPU) ------------------------------------------------------------- */
int64_t swec_expected_shard_size(int64_t dat_size, int data_shards, int64_t large_block,
                                 int64_t small_block);
typedef struct swec_interval {
    int32_t block_index;
    int32_t is_large_block;
    int64_t inner_block_offset;
    int64_t size;
    int32_t large_block_rows_count;
    int32_t reserved;
} swec_interval;
/* Returns the number of intervals written, or SWEC_ERR_INVALID_ARG if cap is too small.         */
int swec_locate_data(int64_t large_block, int64_t small_block, int64_t shard_dat_size,
                     int64_t offset, int64_t size, int data_shards, swec_interval *out, int cap);
void swec_interval_to_shard(const swec_interval *iv, int64_t large_block, int64_t small_block,
                            int data_shards, int *shard_id, int64_t *shard_offset);

/* ---- pinned host memory for callers that stage their own buffers ---------------------------- */
void *swec_alloc_pinned(size_t bytes);
/* Same, bound to the NUMA node the GPU `device` is attached to (full PCIe rate on 2-socket hosts). */
void *swec_alloc_pinned_for_device(int device, size_t bytes);
void swec_free_pinned(void *p);

/* ---- measurement helpers (synthetic volumes, device digests) -------------------------------- */
/* byte b of the stream = byte (b%8) of splitmix64(seed + (b/8 + 1)·0x9E3779B97F4A7C15).        */
int swec_synth_fill_device(int device, void *dst, uint64_t byte_offset, size_t bytes,
                           uint64_t seed, void *stream);
/* 64-bit order-sensitive digest of a device buffer; synchronises `stream`.                      */
int swec_digest_device(int device, const void *src, size_t bytes, uint64_t *digest, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SWEC_H */
