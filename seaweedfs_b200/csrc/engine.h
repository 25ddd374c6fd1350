// seaweedfs_b200/csrc/engine.h — the encoder object behind the C ABI (include/swec.h).
//
// Mirrors what SeaweedFS holds as a reedsolomon.Encoder (ec_context.go:34-36): an immutable
// generator matrix plus, here, the device-side state needed to apply matrices on a B200:
// a stream, cached multiply tables / specialised kernels per matrix, and a small ring of pinned
// + device staging buffers for calls that arrive with host memory.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/swec.h"
#include "gf256.h"
#include "kernels.h"

namespace swec {

void set_last_error(const std::string& msg);
const char* last_error();
int fail(int status, const std::string& msg);
int cuda_fail(cudaError_t e, const char* what);

#define SWEC_CUDA(call)                                     \
    do {                                                    \
        cudaError_t e__ = (call);                           \
        if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
    } while (0)

// pinned host memory on the NUMA node of `device` (plain cudaHostAlloc when that is unknown)
void* pinned_alloc(int device, size_t bytes);
void pinned_free(void* p);
int device_numa_node(int device);

// ecShardConfig.{dataShards,parityShards} of a .vif file (ec_files.cc); false when absent/unreadable
bool read_vif_ratio(const std::string& path, int* ds, int* ps);
// "file_direct_io" (SWEC_FILE_DIRECT): bit 0 = O_DIRECT reads of the .dat / shard inputs straight into the pinned
// ring, bit 1 = O_DIRECT writes of the shard outputs — the page cache is bypassed both ways (disk-backed volumes only;
// files that refuse O_DIRECT, tmpfs for one, and unaligned pieces silently take the buffered descriptor)
extern std::atomic<long> g_opt_file_direct_io;
void file_pipeline_trim();  // ec_files.cc: release staging rings parked between file-level calls

// device-resident multiply tables of one R×K matrix (R ≤ 4)
struct DeviceTables {
    u32* compact = nullptr;     // [K][2][16]
    u32* replicated = nullptr;  // [K][2][16][32]
};

struct JitKernel;  // jit.cc

struct StagingSlot {
    uint8_t* host = nullptr;      // pinned, (k+2m)*chunk
    uint8_t* host_dev = nullptr;  // the same memory as the GPU addresses it (mapped pinned memory), or nullptr
    uint8_t* dev = nullptr;       // (k+2m)*chunk
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
    bool busy = false;
};

struct Layout {  // how stream i / column x maps to memory; see SwecApplyParams
    bool blocked = false;
    uint64_t block_bytes = 0;
};

struct swec_encoder_impl {
    int k = 0, m = 0, device = -1;
    Matrix gen;
    bool rs10x4 = false;

    std::mutex mu;  // guards everything below
    cudaStream_t stream = nullptr;
    std::map<std::vector<uint8_t>, DeviceTables> tables;  // key: R, K, coefficients
    std::map<std::vector<uint8_t>, std::shared_ptr<JitKernel>> jit;  // same key
    std::vector<StagingSlot> slots;
    size_t slot_chunk = 0;
    // file pipelines never stall on a kernel compile: a cold matrix is served by the table kernel (100x faster
    // than the I/O around it) while the specialised kernel is built in the background and picked up when ready
    bool never_wait_for_jit = false;
    uint8_t* tail_scratch = nullptr;  // k*small zero-padded last row
    size_t tail_scratch_bytes = 0;

    ~swec_encoder_impl();
    int ensure_device();  // cudaSetDevice + lazily create stream
    int ensure_slots(size_t chunk);

    // out[r][x] = XOR_i rows[r][i] ⊗ in[i][x] on device memory, asynchronous on s.
    int apply(const Matrix& rows, const uint8_t* const* in, uint8_t* const* out, size_t n,
              const Layout& layout, cudaStream_t s);
    int get_tables(const Matrix& rows4, DeviceTables* out, cudaStream_t s);
};

// true when run-time specialised kernels can be had: NVRTC loaded (SWEC_NO_JIT unset), or the on-disk cubin cache is on
bool jit_available();
unsigned long long jit_compile_count();   // NVRTC compiles done by this process
unsigned long long jit_disk_hit_count();  // kernels loaded from the on-disk cubin cache instead
// Specialised Horner kernel for `rows` on the current device (compiled once per matrix/process).
int jit_get(swec_encoder_impl* enc, const Matrix& rows, std::shared_ptr<JitKernel>* out, bool wait = true, bool hot = false);
int jit_debug_compile(const Matrix& rows, size_t* cubin_bytes, int* xtime_steps, int* xor_ops);
void jit_shutdown();  // stop the background compiler (idempotent); inline compiles keep working
bool jit_cached(swec_encoder_impl* enc, const Matrix& rows);
cudaError_t jit_launch(const JitKernel& k, const SwecApplyParams& p, bool blocked, cudaStream_t s);

}  // namespace swec

struct swec_encoder : swec::swec_encoder_impl {};
