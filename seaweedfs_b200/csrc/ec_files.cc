// seaweedfs_b200/csrc/ec_files.cc — file-level entry points: the B200 twins of
//   generateEcFiles / encodeDatFile / encodeData / encodeDataOneBatch
//       (weed/storage/erasure_coding/ec_encoder.go:110-128, 202-222, 248-278, 280-321)
//   generateMissingEcFiles / rebuildEcFiles / findShardFile      (ec_encoder.go:131-200, 323-377)
//   WriteDatFile                                                (ec_decoder.go:176-223)
// Same files, same bytes, same error points; different schedule.  The reference runs
// read → Encode → write serially in 256 KiB batches on one goroutine.  Here a reader stages multi-MiB
// stripes into pinned slots, the GPU turns each slot round on its own stream (H2D, kernel, D2H)
// and a writer thread drains finished slots with pwrite at explicit shard offsets, so disk reads,
// PCIe, the kernel and disk writes all overlap.  Batch size is result-neutral (the code is
// column-wise), so buffer_size is only validated the way the reference does.
#include <errno.h>
#include <fcntl.h>
#include <libgen.h>
#include <linux/falloc.h>
#include <sys/stat.h>
#include <sys/vfs.h>
#include <unistd.h>

#include <algorithm>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>
#include <memory>
#include <thread>
#include <time.h>

#include "engine.h"
#include "io_pool.h"
#include "mini_json.h"

namespace swec {
namespace {

size_t env_sz(const char* name, size_t dflt) {
    const char* e = getenv(name);
    const long long v = e ? atoll(e) : 0;
    return v > 0 ? size_t(v) : dflt;
}

int io_fail(const std::string& what) { return fail(SWEC_ERR_IO, what + ": " + strerror(errno)); }

std::string shard_ext(int idx) {  // ToExt, ec_encoder.go:106-108
    char b[16];
    snprintf(b, sizeof b, ".ec%02d", idx);
    return b;
}

// Reserving extents pays on disk filesystems (once instead of 14 files growing 8 MiB at a time) and costs on tmpfs,
// where it zero-fills every page that the writers overwrite a moment later.  The reservation must NOT change the
// visible file size: DiskLocation.validateEcVolume (disk_location_ec.go:455-530) and rebuildEcFiles' equal-length
// check recognise an interrupted encode by its short shards, so a killed run has to leave short files, not
// full-size files of zeros — hence FALLOC_FL_KEEP_SIZE (reserve_extents), never posix_fallocate.
bool worth_preallocating(int fd) {
    if (getenv("SWEC_NO_FALLOCATE")) return false;
    struct statfs fs;
    if (fstatfs(fd, &fs) != 0) return true;
    return fs.f_type != 0x01021994 /* TMPFS_MAGIC */ && fs.f_type != 0x858458f6 /* RAMFS_MAGIC */;
}

void reserve_extents(int fd, int64_t size) {  // best effort; the file's length stays what has been written
    if (size > 0 && fallocate(fd, FALLOC_FL_KEEP_SIZE, 0, off_t(size)) != 0) errno = 0;
}

// a second descriptor on the same file with O_DIRECT, or -1 (tmpfs and friends refuse it; so may the caller's option)
int open_direct(const std::string& path, int flags, bool wanted) {
    if (!wanted) return -1;
    const int fd = open(path.c_str(), flags | O_DIRECT);
    if (fd < 0) errno = 0;
    return fd;
}
inline bool direct_ok(int dfd, int64_t off, size_t len, const void* buf) {
    return dfd >= 0 && ((uint64_t(off) | uint64_t(len) | reinterpret_cast<uintptr_t>(buf)) & 4095) == 0;
}

struct FdSet {
    std::vector<int> fds;
    ~FdSet() {
        for (int fd : fds)
            if (fd >= 0) close(fd);
    }
};

// fill bytes [dst_off, dst_off+len) of the slot's stream `stream` from fd@off (zero past EOF)
// dfd: the same file opened O_DIRECT (-1 = none): used when offset, length and buffer are all 4 KiB aligned
struct ReadOp { int stream, fd; int64_t off; size_t dst_off, len; int dfd = -1; };
struct WriteOp { int stream, fd; int64_t off; int dfd = -1; };  // drain stream `stream` of the slot to fd@off
struct Item {
    size_t len = 0;
    std::vector<ReadOp> reads;
    std::vector<WriteOp> writes;
};

struct Slot {
    uint8_t* host = nullptr;
    uint8_t* dev = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
    Item item;
};


// ONE I/O pool per process, shared by every file pipeline: concurrent volumes (the shell runs up to 10 at once) must not
// multiply the thread count — 4 pipelines x 16 threads on a 16-core cgroup quota ran at 8.5 GB/s where one volume alone
// does 21 (profiles/r02e_files_multi_1gpu_4x4GiB.jsonl).  Size: the CPU time the process may actually use (cgroup quota,
// else the hardware concurrency), between 4 and 64; SWEC_IO_THREADS overrides.  Leaked on purpose, like the other
// process-wide helpers: threads must not be joined from static destructors.
size_t usable_cpus() {
    size_t n = std::max(1u, std::thread::hardware_concurrency());
    if (FILE* f = fopen("/sys/fs/cgroup/cpu.max", "r")) {  // cgroup v2: "<quota|max> <period>"
        char q[32] = {0};
        long long period = 0;
        if (fscanf(f, "%31s %lld", q, &period) == 2 && strcmp(q, "max") != 0 && period > 0)
            n = std::min<size_t>(n, size_t(std::max<long long>(1, (atoll(q) + period - 1) / period)));
        fclose(f);
    } else {
        long long quota = -1, period = 0;
        if (FILE* g = fopen("/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "r")) {
            if (fscanf(g, "%lld", &quota) != 1) quota = -1;
            fclose(g);
        }
        if (FILE* g = fopen("/sys/fs/cgroup/cpu/cpu.cfs_period_us", "r")) {
            if (fscanf(g, "%lld", &period) != 1) period = 0;
            fclose(g);
        }
        if (quota > 0 && period > 0) n = std::min<size_t>(n, size_t(std::max<long long>(1, (quota + period - 1) / period)));
    }
    return n;
}
IoPool& file_io_pool() {
    static IoPool* pool = new IoPool(env_sz("SWEC_IO_THREADS", std::min<size_t>(64, std::max<size_t>(4, usable_cpus()))));
    return *pool;
}

// Staging rings outlive a call: pinning (mmap + mbind + cudaHostRegister) and un-pinning 3 x 14 x 8 MiB costs
// 0.1-2 s per call (profiles/r01z_files_*), as much as the pipeline itself spends on an 8 GiB volume.  A volume
// server encodes volume after volume, so finished pipelines park their ring here (per device and size, a few at
// most) and the next call picks it up.  swec_shutdown() releases them.
struct SlotSet {
    int device = -1;
    size_t bytes_per_slot = 0;
    std::vector<Slot> slots;
};
std::mutex& slotset_mu() {
    static std::mutex* m = new std::mutex;
    return *m;
}
std::vector<SlotSet>& slotset_cache() {
    static std::vector<SlotSet>* c = new std::vector<SlotSet>;  // leaked on purpose: no CUDA calls in static destructors
    return *c;
}
constexpr size_t kMaxCachedSlotSets = 4;

void free_slots(int device, std::vector<Slot>& slots) {
    if (cudaSetDevice(device) != cudaSuccess) cudaGetLastError();
    for (auto& s : slots) {
        if (s.stream) cudaStreamSynchronize(s.stream);
        if (s.host) pinned_free(s.host);
        if (s.dev) cudaFree(s.dev);
        if (s.done) cudaEventDestroy(s.done);
        if (s.stream) cudaStreamDestroy(s.stream);
    }
    slots.clear();
}

// wall-clock breakdown of one pipeline run, printed to stderr as JSON when SWEC_PIPE_STATS is set
struct PipeStats {
    double setup = 0, prealloc = 0, wait_slot = 0, read = 0, enqueue = 0, wait_gpu = 0, write = 0, total = 0;
    int items = 0;
    static double now() {
        struct timespec t;
        clock_gettime(CLOCK_MONOTONIC, &t);
        return double(t.tv_sec) + double(t.tv_nsec) * 1e-9;
    }
};

// K input streams + R computed streams per slot, pitch = chunk bytes.
class FilePipeline {
  public:
    PipeStats stats;
    // verify = true: streams [K, K+R) are the parity bytes read from disk; the computed parity goes to
    // streams [K+R, K+2R) and is only compared on the device (no D2H, no writes).
    FilePipeline(swec_encoder* enc, const Matrix& rows, size_t chunk, bool verify = false)
        : enc_(enc), rows_(rows), chunk_(chunk), verify_(verify) {}
    ~FilePipeline() { shutdown(); }

    int start() {
        const double t_start = PipeStats::now();
        struct SetupTimer {
            PipeStats& st;
            double t0;
            ~SetupTimer() { st.setup += PipeStats::now() - t0; st.total -= 0; }
        } setup_timer{stats, t_start};
        t_begin_ = t_start;
        enc_->never_wait_for_jit = !getenv("SWEC_FILE_JIT_WAIT");
        int rc = enc_->ensure_device();
        if (rc) return rc;
        const size_t nslots = env_sz("SWEC_STAGE_SLOTS", 3);
        const size_t streams = size_t(rows_.cols + rows_.rows * (verify_ ? 2 : 1));
        if (verify_) {
            SWEC_CUDA(cudaMalloc(reinterpret_cast<void**>(&dev_bad_), sizeof(unsigned long long) * size_t(rows_.rows)));
            SWEC_CUDA(cudaMemset(dev_bad_, 0, sizeof(unsigned long long) * size_t(rows_.rows)));
        }
        // one size for every kind of pipeline of this code on this device (generate: k+m streams, rebuild: k + missing,
        // verify: k+2m), so that a parked ring fits whichever call comes next
        slot_bytes_ = std::max(streams, size_t(enc_->k + 2 * enc_->m)) * chunk_;
        {
            std::lock_guard<std::mutex> lk(slotset_mu());
            auto& cache = slotset_cache();
            for (size_t i = 0; i < cache.size(); i++)
                if (cache[i].device == enc_->device && cache[i].bytes_per_slot == slot_bytes_ && cache[i].slots.size() == nslots) {
                    slots_ = std::move(cache[i].slots);
                    cache.erase(cache.begin() + long(i));
                    break;
                }
        }
        if (slots_.empty()) {
            slots_.resize(nslots);
            for (auto& s : slots_) {
                s.host = static_cast<uint8_t*>(pinned_alloc(enc_->device, slot_bytes_));
                cudaError_t e = s.host ? cudaSuccess : cudaErrorMemoryAllocation;
                if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&s.dev), slot_bytes_);
                if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking);
                if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming);
                if (e != cudaSuccess) {
                    const int frc = s.host ? cuda_fail(e, "allocating the staging ring") : fail(SWEC_ERR_NOMEM, "cannot allocate pinned staging memory");
                    free_slots(enc_->device, slots_);
                    return frc;
                }
            }
        }
        for (auto& s : slots_) {
            s.item = Item{};
            free_.push_back(&s);
        }
        io_ = &file_io_pool();
        writer_ = std::thread([this] { writer_loop(); });
        started_ = true;
        return SWEC_OK;
    }

    // blocking: read the item's inputs, queue the GPU work, hand the slot to the writer
    int submit(Item&& item) {
        Slot* s = nullptr;
        double t0 = PipeStats::now();
        {
            std::unique_lock<std::mutex> lk(mu_);
            cv_.wait(lk, [&] { return !free_.empty() || error_; });
            if (error_) return error_;
            s = free_.front();
            free_.pop_front();
        }
        double t1 = PipeStats::now();
        stats.wait_slot += t1 - t0;
        stats.items++;
        const int K = rows_.cols, R = rows_.rows;
        const size_t len = item.len;
        // Every pread is cut into pieces of at most io_piece_ bytes so that ONE stripe keeps the whole I/O pool busy
        // (k preads of 8 MiB are only k tasks; a thread copies 2-3 GB/s out of the page cache into cache-cold memory).
        struct Piece { int op; size_t off, len; };
        std::vector<Piece> rpieces;
        for (size_t i = 0; i < item.reads.size(); i++)
            for (size_t o = 0; o < item.reads[i].len; o += io_piece_)
                rpieces.push_back({int(i), o, std::min(io_piece_, item.reads[i].len - o)});
        const std::function<int(int)> read_one = [&](int idx) -> int {
            const Piece& pc = rpieces[size_t(idx)];
            const ReadOp& r = item.reads[size_t(pc.op)];
            uint8_t* dst = s->host + size_t(r.stream) * chunk_ + r.dst_off + pc.off;
            const int64_t off = r.off + int64_t(pc.off);
            const size_t len = pc.len;
            // O_DIRECT: the device DMAs straight into the pinned slot; short counts only happen at EOF, where the
            // remainder (unaligned now) continues on the buffered descriptor
            const int fd0 = direct_ok(r.dfd, off, len, dst) ? r.dfd : r.fd;
            size_t got = 0;
            while (got < len) {
                const int fd = (fd0 == r.dfd && direct_ok(r.dfd, off + int64_t(got), len - got, dst + got)) ? r.dfd : r.fd;
                const ssize_t n = pread(fd, dst + got, len - got, off_t(off + int64_t(got)));
                if (n < 0) {
                    if (errno == EINTR) continue;
                    const int rc = io_fail("pread");
                    note_error(last_error());
                    return rc;
                }
                if (n == 0) break;  // EOF: the rest reads as zero (ec_encoder.go:258-262)
                got += size_t(n);
            }
            if (got < len) memset(dst + got, 0, len - got);
            return SWEC_OK;
        };
        if (const int rrc = io_->parallel_for(int(rpieces.size()), read_one)) {
            set_last_error(noted_error());
            return set_error(rrc, s);
        }
        t0 = PipeStats::now();
        stats.read += t0 - t1;
        if (cudaSetDevice(enc_->device) != cudaSuccess) return set_error(fail(SWEC_ERR_CUDA, "cudaSetDevice"), s);
        cudaError_t e = cudaSuccess;
        const uint8_t* din[SWEC_MAX_SHARDS];
        uint8_t* dout[SWEC_MAX_SHARDS];
        for (int i = 0; i < K; i++) din[i] = s->dev + size_t(i) * chunk_;
        const int nin = K + (verify_ ? R : 0);
        if (len == chunk_) {  // full slot: the input streams are contiguous — one DMA
            e = cudaMemcpyAsync(s->dev, s->host, size_t(nin) * chunk_, cudaMemcpyHostToDevice, s->stream);
        } else {
            for (int i = 0; i < nin && e == cudaSuccess; i++)
                e = cudaMemcpyAsync(s->dev + size_t(i) * chunk_, s->host + size_t(i) * chunk_, len, cudaMemcpyHostToDevice, s->stream);
        }
        if (e != cudaSuccess) return set_error(cuda_fail(e, "H2D"), s);
        for (int r = 0; r < R; r++) dout[r] = s->dev + size_t(K + (verify_ ? R : 0) + r) * chunk_;
        int rc;
        {
            std::lock_guard<std::mutex> lk(enc_->mu);
            rc = enc_->apply(rows_, din, dout, len, Layout{}, s->stream);
        }
        if (rc) return set_error(rc, s);
        if (verify_) {
            for (int r = 0; r < R && e == cudaSuccess; r++)
                e = launch_compare(dout[r], s->dev + size_t(K + r) * chunk_, len, dev_bad_ + r, s->stream);
        } else if (len == chunk_) {
            e = cudaMemcpyAsync(s->host + size_t(K) * chunk_, dout[0], size_t(R) * chunk_, cudaMemcpyDeviceToHost, s->stream);
        } else {
            for (int r = 0; r < R && e == cudaSuccess; r++)
                e = cudaMemcpyAsync(s->host + size_t(K + r) * chunk_, dout[r], len, cudaMemcpyDeviceToHost, s->stream);
        }
        if (e == cudaSuccess) e = cudaEventRecord(s->done, s->stream);
        if (e != cudaSuccess) return set_error(cuda_fail(e, "D2H"), s);
        s->item = std::move(item);
        {
            std::lock_guard<std::mutex> lk(mu_);
            inflight_.push_back(s);
        }
        cv_.notify_all();
        stats.enqueue += PipeStats::now() - t0;
        return SWEC_OK;
    }

    int parallel_for(int n, const std::function<int(int)>& fn) { return io_->parallel_for(n, fn); }

    // wait for everything queued so far; returns the first error
    int finish() {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return (inflight_.empty() && free_.size() == slots_.size()) || error_; });
        return error_;
    }

    void report(const char* what) {
        if (!getenv("SWEC_PIPE_STATS")) return;
        stats.total = PipeStats::now() - t_begin_;
        fprintf(stderr,
                "{\"pipe\": \"%s\", \"items\": %d, \"chunk\": %zu, \"total_s\": %.3f, \"setup_s\": %.3f, \"prealloc_s\": %.3f, "
                "\"reader\": {\"wait_slot_s\": %.3f, \"read_s\": %.3f, \"enqueue_s\": %.3f}, "
                "\"writer\": {\"wait_gpu_s\": %.3f, \"write_s\": %.3f}}\n",
                what, stats.items, chunk_, stats.total, stats.setup, stats.prealloc, stats.wait_slot, stats.read, stats.enqueue,
                stats.wait_gpu, stats.write);
    }

    void shutdown() {
        if (!started_) return;
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
        }
        cv_.notify_all();
        if (writer_.joinable()) writer_.join();
        io_ = nullptr;
        cudaSetDevice(enc_->device);
        bool healthy = error_ == 0;
        for (auto& s : slots_)
            if (s.stream && cudaStreamSynchronize(s.stream) != cudaSuccess) {
                cudaGetLastError();
                healthy = false;
            }
        free_.clear();
        inflight_.clear();
        if (healthy && !slots_.empty() && !getenv("SWEC_NO_RING_CACHE")) {  // park the ring for the next call
            std::lock_guard<std::mutex> lk(slotset_mu());
            auto& cache = slotset_cache();
            if (cache.size() < kMaxCachedSlotSets) {
                SlotSet set;
                set.device = enc_->device;
                set.bytes_per_slot = slot_bytes_;
                set.slots = std::move(slots_);
                cache.push_back(std::move(set));
                slots_.clear();
            }
        }
        if (!slots_.empty()) free_slots(enc_->device, slots_);
        if (dev_bad_) cudaFree(dev_bad_);
        dev_bad_ = nullptr;
        started_ = false;
    }

    // verify mode, after finish(): mismatching 16-byte vectors per parity row
    int mismatches(unsigned long long* out) {
        if (!dev_bad_) return fail(SWEC_ERR_INVALID_ARG, "not a verify pipeline");
        SWEC_CUDA(cudaMemcpy(out, dev_bad_, sizeof(unsigned long long) * size_t(rows_.rows), cudaMemcpyDeviceToHost));
        return SWEC_OK;
    }

  private:
    // I/O runs on pool threads whose thread-local error text the submitting thread cannot see
    void note_error(const char* msg) {
        std::lock_guard<std::mutex> lk(note_mu_);
        if (noted_.empty()) noted_ = msg;
    }
    std::string noted_error() {
        std::lock_guard<std::mutex> lk(note_mu_);
        return noted_;
    }
    std::mutex note_mu_;
    std::string noted_;

    int set_error(int rc, Slot* s) {
        std::lock_guard<std::mutex> lk(mu_);
        if (!error_) {
            error_ = rc;
            error_msg_ = last_error();
        }
        if (s) free_.push_back(s);
        cv_.notify_all();
        return rc;
    }

    void writer_loop() {
        cudaSetDevice(enc_->device);
        for (;;) {
            Slot* s = nullptr;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return !inflight_.empty() || stop_; });
                if (inflight_.empty()) return;
                s = inflight_.front();
            }
            int rc = SWEC_OK;
            const double tw0 = PipeStats::now();
            if (cudaEventSynchronize(s->done) != cudaSuccess) rc = fail(SWEC_ERR_CUDA, "cudaEventSynchronize failed in the shard writer");
            const double tw1 = PipeStats::now();
            stats.wait_gpu += tw1 - tw0;   // writer thread only
            if (!rc) {
                struct Piece { int op; size_t off, len; };
                std::vector<Piece> wpieces;
                // one task per shard file: buffered writes take the inode's lock exclusively, so pieces of the same
                // file would only queue behind each other (SWEC_FILE_WRITE_PIECE splits them anyway, for O_DIRECT devices)
                const size_t wpiece = std::max<size_t>(4096, env_sz("SWEC_FILE_WRITE_PIECE", s->item.len ? s->item.len : 4096) & ~size_t(4095));
                for (size_t i = 0; i < s->item.writes.size(); i++)
                    for (size_t o = 0; o < s->item.len; o += wpiece)
                        wpieces.push_back({int(i), o, std::min(wpiece, s->item.len - o)});
                const std::function<int(int)> write_one = [&](int idx) -> int {
                    const Piece& pc = wpieces[size_t(idx)];
                    const WriteOp& w = s->item.writes[size_t(pc.op)];
                    const uint8_t* src = s->host + size_t(w.stream) * chunk_ + pc.off;
                    const int64_t off = w.off + int64_t(pc.off);
                    size_t put = 0;
                    while (put < pc.len) {
                        const int fd = direct_ok(w.dfd, off + int64_t(put), pc.len - put, src + put) ? w.dfd : w.fd;
                        const ssize_t n = pwrite(fd, src + put, pc.len - put, off_t(off + int64_t(put)));
                        if (n < 0) {
                            if (errno == EINTR) continue;
                            const int rc2 = io_fail("pwrite");
                            note_error(last_error());
                            return rc2;
                        }
                        put += size_t(n);
                    }
                    return SWEC_OK;
                };
                rc = io_->parallel_for(int(wpieces.size()), write_one);
                if (rc) set_last_error(noted_error());
                stats.write += PipeStats::now() - tw1;
            }
            {
                std::lock_guard<std::mutex> lk(mu_);
                inflight_.pop_front();
                free_.push_back(s);
                if (rc && !error_) {
                    error_ = rc;
                    error_msg_ = last_error();
                }
            }
            cv_.notify_all();
        }
    }

    swec_encoder* enc_;
    Matrix rows_;
    size_t chunk_;
    const size_t io_piece_ = std::max<size_t>(4096, env_sz("SWEC_FILE_IO_PIECE", size_t(2) << 20) & ~size_t(4095));
    double t_begin_ = 0;
    bool verify_ = false;
    unsigned long long* dev_bad_ = nullptr;
    std::vector<Slot> slots_;
    size_t slot_bytes_ = 0;
    std::deque<Slot*> free_, inflight_;
    std::mutex mu_;
    std::condition_variable cv_;
    std::thread writer_;
    IoPool* io_ = nullptr;  // the process-wide pool (not owned)
    int error_ = 0;
    std::string error_msg_;
    bool stop_ = false, started_ = false;

  public:
    const std::string& error_message() const { return error_msg_; }
};

}  // namespace

// .vif is protobuf-JSON (weed/storage/volume_info/volume_info.go:73-95); we only need
// ecShardConfig.{dataShards,parityShards} (weed/pb/volume_server.proto:561-577).
bool read_vif_ratio(const std::string& path, int* ds, int* ps) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    std::string txt;
    char buf[4096];
    size_t n;
    while ((n = fread(buf, 1, sizeof buf, f)) > 0) txt.append(buf, n);
    fclose(f);
    int64_t a = 0, b = 0;
    if (!mini_json::nested_int(txt, "ecShardConfig", "ec_shard_config", "dataShards", "data_shards", &a) ||
        !mini_json::nested_int(txt, "ecShardConfig", "ec_shard_config", "parityShards", "parity_shards", &b))
        return false;
    if (a < 0 || b < 0 || a > 255 || b > 255) return false;
    *ds = int(a);
    *ps = int(b);
    return true;
}

namespace {

bool file_exists(const std::string& p) {
    struct stat st;
    return stat(p.c_str(), &st) == 0 && !S_ISDIR(st.st_mode);
}

}  // namespace

void file_pipeline_trim() {  // swec_shutdown(): release parked staging rings
    std::vector<SlotSet> sets;
    {
        std::lock_guard<std::mutex> lk(slotset_mu());
        sets.swap(slotset_cache());
    }
    for (auto& set : sets) free_slots(set.device, set.slots);
}

}  // namespace swec

using namespace swec;

extern "C" {

int swec_generate_ec_files(const char* base, int64_t buffer_size, int64_t large, int64_t small, int k, int m,
                           int device) {
    if (!base) return fail(SWEC_ERR_INVALID_ARG, "base_file_name is NULL");
    const double t_call = PipeStats::now();
    // encodeData: "unexpected zero buffer size" / "unexpected block size %d buffer size %d" (ec_encoder.go:204-212)
    if (buffer_size <= 0 || large <= 0 || small <= 0 || large % buffer_size || small % buffer_size)
        return fail(SWEC_ERR_INVALID_ARG, "block sizes must be positive multiples of buffer_size");
    swec_encoder* enc = nullptr;
    int rc = swec_encoder_new(k, m, device, &enc);
    if (rc) return rc;
    std::unique_ptr<swec_encoder, void (*)(swec_encoder*)> guard(enc, swec_encoder_free);

    const std::string b(base);
    FdSet fds;
    const int dat = open((b + ".dat").c_str(), O_RDONLY);
    if (dat < 0) return io_fail("failed to open dat file " + b + ".dat");
    fds.fds.push_back(dat);
    struct stat st;
    if (fstat(dat, &st) != 0) return io_fail("failed to stat dat file");
    const int total = k + m;
    const long direct = g_opt_file_direct_io.load();
    const int dat_d = open_direct(b + ".dat", O_RDONLY, direct & 1);
    if (dat_d >= 0) fds.fds.push_back(dat_d);
    std::vector<int> outs(static_cast<size_t>(total), -1), outs_d(static_cast<size_t>(total), -1);
    {
        // openEcFiles (ec_encoder.go:224-238): O_TRUNC|O_CREAT|O_WRONLY 0644 for every shard — all at once.  Truncating
        // a shard file left by an earlier encode frees its pages one file after the other when done in a loop: 2.2 s
        // for the 11 GiB of shards of an 8 GiB volume on tmpfs, five times the encode itself
        // (write_ec_files_over_existing_shards_GBps in profiles/r02a_bench_n1.json: 3.95 vs 18.7 GB/s).
        std::vector<int> err(static_cast<size_t>(total), 0);
        std::vector<std::thread> openers;
        for (int i = 0; i < total; i++)
            openers.emplace_back([&, i] {
                outs[size_t(i)] = open((b + shard_ext(i)).c_str(), O_TRUNC | O_CREAT | O_WRONLY, 0644);
                if (outs[size_t(i)] < 0) err[size_t(i)] = errno;
                else outs_d[size_t(i)] = open_direct(b + shard_ext(i), O_WRONLY, direct & 2);
            });
        for (auto& t : openers) t.join();
        for (int i = 0; i < total; i++) {
            if (outs[size_t(i)] >= 0) fds.fds.push_back(outs[size_t(i)]);
            if (outs_d[size_t(i)] >= 0) fds.fds.push_back(outs_d[size_t(i)]);
        }
        for (int i = 0; i < total; i++)
            if (outs[size_t(i)] < 0) {
                errno = err[size_t(i)];
                return io_fail("failed to open file " + b + shard_ext(i));
            }
    }

    Matrix rows(m, k);
    memcpy(rows.v.data(), enc->gen.row(k), rows.v.size());
    const int64_t max_chunk = int64_t(env_sz("SWEC_FILE_CHUNK", size_t(8) << 20));
    const size_t chunk = size_t(std::min<int64_t>(max_chunk, std::max(large, small)) + 255) & ~size_t(255);
    const double t_opened = PipeStats::now();
    FilePipeline pipe(enc, rows, chunk);
    if ((rc = pipe.start())) return rc;

    // The final shard size is known up front: reserve its extents (visible length unchanged) — one I/O thread per
    // file — so the writers fill pages/extents that already exist instead of
    // growing 14 files 8 MiB at a time under the filesystem's allocation lock (best effort).
    if (worth_preallocating(outs[0])) {
        const double tp = PipeStats::now();
        const int64_t shard_size = swec_expected_shard_size(st.st_size, k, large, small);
        if (shard_size > 0)
            pipe.parallel_for(total, [&](int i) -> int {
                reserve_extents(outs[size_t(i)], shard_size);
                return 0;
            });
        pipe.stats.prealloc += PipeStats::now() - tp;
    }

    int64_t remaining = st.st_size, processed = 0, shard_off = 0;
    const int64_t large_row = large * k, small_row = small * k;
    auto encode_row = [&](int64_t block) -> int {  // encodeData on one row of k blocks, chunk by chunk
        for (int64_t o = 0; o < block; o += int64_t(chunk)) {
            Item it;
            it.len = size_t(std::min<int64_t>(int64_t(chunk), block - o));
            for (int i = 0; i < k; i++) it.reads.push_back({i, dat, processed + block * i + o, 0, it.len, dat_d});
            for (int i = 0; i < total; i++) it.writes.push_back({i, outs[size_t(i)], shard_off + o, outs_d[size_t(i)]});
            const int r = pipe.submit(std::move(it));
            if (r) return r;
        }
        shard_off += block;
        return SWEC_OK;
    };
    while (rc == SWEC_OK && remaining >= large_row) {  // ec_encoder.go:304-311
        rc = encode_row(large);
        remaining -= large_row;
        processed += large_row;
    }
    while (rc == SWEC_OK && remaining > 0 && small > int64_t(chunk)) {  // small blocks bigger than a slot: row by row
        rc = encode_row(small);
        remaining -= small_row;
        processed += small_row;
    }
    // Small rows (ec_encoder.go:312-319) are tiny (10 x 1 MiB): many of them share one slot.  Row j of the
    // batch is one contiguous k*small run of the .dat whose k blocks scatter to offset j*small of the k input
    // streams, so every shard still receives ONE contiguous write per item.  A default 30,000 MiB volume is
    // 2 large rows + 952 small ones — a third of its bytes take this path.
    const int64_t rows_per_item = std::max<int64_t>(1, int64_t(chunk) / small);
    while (rc == SWEC_OK && remaining > 0) {
        const int64_t rows_left = (remaining + small_row - 1) / small_row;
        const int64_t n = std::min(rows_per_item, rows_left);
        Item it;
        it.len = size_t(n * small);
        for (int64_t j = 0; j < n; j++)
            for (int i = 0; i < k; i++)
                it.reads.push_back({i, dat, processed + j * small_row + int64_t(i) * small, size_t(j * small), size_t(small), dat_d});
        for (int i = 0; i < total; i++) it.writes.push_back({i, outs[size_t(i)], shard_off, outs_d[size_t(i)]});
        rc = pipe.submit(std::move(it));
        shard_off += n * small;
        remaining -= n * small_row;
        processed += n * small_row;
    }
    const int rc2 = pipe.finish();
    if (rc == SWEC_OK) rc = rc2;
    pipe.report("generate_ec_files");
    const double t_piped = PipeStats::now();
    const std::string msg = pipe.error_message();
    pipe.shutdown();
    if (getenv("SWEC_PIPE_STATS"))
        fprintf(stderr, "{\"call\": \"generate_ec_files\", \"dat_bytes\": %lld, \"open_and_truncate_s\": %.3f, \"pipeline_s\": %.3f, "
                        "\"teardown_s\": %.3f}\n",
                (long long)st.st_size, t_opened - t_call, t_piped - t_opened, PipeStats::now() - t_piped);
    if (rc && !msg.empty()) set_last_error(msg);
    return rc;
}

int swec_write_ec_files(const char* base, int device) {
    // WriteEcFilesWithContext: 256 KiB buffers, 1 GiB / 1 MiB blocks, 10+4 (ec_encoder.go:61-69)
    return swec_generate_ec_files(base, 256 * 1024, int64_t(1) << 30, int64_t(1) << 20, 10, 4, device);
}

int swec_rebuild_ec_files(const char* base, const char* const* dirs, int ndirs, int k, int m, int device,
                          uint32_t* rebuilt, int* n_rebuilt) {
    if (!base || !rebuilt || !n_rebuilt || (ndirs > 0 && !dirs)) return fail(SWEC_ERR_INVALID_ARG, "NULL argument");
    *n_rebuilt = 0;
    const std::string b(base);
    if (k == 0) {  // RebuildEcFiles: ratio from .vif when valid, else default (ec_encoder.go:76-95)
        int ds = 0, ps = 0;
        if (read_vif_ratio(b + ".vif", &ds, &ps) && ds > 0 && ps > 0 && ds + ps <= SWEC_MAX_SHARDS) {
            k = ds;
            m = ps;
        } else {
            k = 10;
            m = 4;
        }
    }
    swec_encoder* enc = nullptr;
    int rc = swec_encoder_new(k, m, device, &enc);
    if (rc) return rc;
    std::unique_ptr<swec_encoder, void (*)(swec_encoder*)> guard(enc, swec_encoder_free);
    const int total = k + m;

    // pass 1: which shards exist (base dir first, then additionalDirs) — ec_encoder.go:131-169
    std::string base_copy(b);
    const std::string base_name = basename(&base_copy[0]);
    FdSet fds;
    const long direct = g_opt_file_direct_io.load();
    std::vector<int> in(static_cast<size_t>(total), -1), in_d(static_cast<size_t>(total), -1), out_d(static_cast<size_t>(total), -1);
    std::vector<uint8_t> present(static_cast<size_t>(total), 0);
    int npresent = 0;
    std::vector<uint32_t> missing;
    for (int i = 0; i < total; i++) {
        std::string path = b + shard_ext(i);
        if (!file_exists(path)) {
            path.clear();
            for (int d = 0; d < ndirs; d++) {
                const std::string cand = std::string(dirs[d]) + "/" + base_name + shard_ext(i);
                if (file_exists(cand)) {
                    path = cand;
                    break;
                }
            }
        }
        if (path.empty()) {
            missing.push_back(uint32_t(i));
            continue;
        }
        const int fd = open(path.c_str(), O_RDONLY);
        if (fd < 0) return io_fail("open " + path);
        fds.fds.push_back(fd);
        in[size_t(i)] = fd;
        in_d[size_t(i)] = open_direct(path, O_RDONLY, direct & 1);
        if (in_d[size_t(i)] >= 0) fds.fds.push_back(in_d[size_t(i)]);
        present[size_t(i)] = 1;
        npresent++;
    }
    if (npresent < k)  // before any output file exists — ec_encoder.go:172-175
        return fail(SWEC_ERR_TOO_FEW_SHARDS, "not enough shards to rebuild " + b + ": found " + std::to_string(npresent) +
                                                 " shards, need at least " + std::to_string(k));
    for (uint32_t id : missing) rebuilt[(*n_rebuilt)++] = id;
    if (missing.empty()) return SWEC_OK;

    // pass 2: create the outputs — ec_encoder.go:182-193.  Whatever goes wrong from here on, the caller gets no
    // shard ids (the reference returns nil ids with the error) and no half-written output survives: a shard file
    // that exists is taken for a present input by the next rebuild (findShardFile), so a partial one must not stay.
    std::vector<int> out(static_cast<size_t>(total), -1);
    struct Undo {
        const std::string& b;
        const std::vector<uint32_t>& ids;
        int* n_rebuilt;
        size_t created = 0;
        bool armed = true;
        ~Undo() {
            if (!armed) return;
            for (size_t i = 0; i < created; i++) unlink((b + shard_ext(int(ids[i]))).c_str());
            *n_rebuilt = 0;
        }
    } undo{b, missing, n_rebuilt};
    for (uint32_t id : missing) {
        const int fd = open((b + shard_ext(int(id))).c_str(), O_TRUNC | O_WRONLY | O_CREAT, 0644);
        if (fd < 0) return io_fail("create " + b + shard_ext(int(id)));
        undo.created++;
        fds.fds.push_back(fd);
        out[id] = fd;
        out_d[id] = open_direct(b + shard_ext(int(id)), O_WRONLY, direct & 2);
        if (out_d[id] >= 0) fds.fds.push_back(out_d[id]);
    }

    // rebuildEcFiles (ec_encoder.go:323-377): every present shard must have the same length; the
    // reference steps in 1 MiB reads and fails at the first short, unequal one.
    int64_t size = -1;
    for (int i = 0; i < total; i++) {
        if (!present[size_t(i)]) continue;
        struct stat st;
        if (fstat(in[size_t(i)], &st) != 0) return io_fail("fstat shard");
        if (size < 0) size = st.st_size;
        else if (size != st.st_size)
            return fail(SWEC_ERR_SHARD_SIZE, "ec shard size expected " + std::to_string(size) + " actual " + std::to_string(st.st_size));
    }
    const int64_t mib = int64_t(1) << 20;
    // quirk kept: a length above 1 MiB that is not a multiple of 1 MiB errors on the last read
    const bool ragged = size > mib && size % mib != 0;
    const int64_t todo = ragged ? size / mib * mib : size;

    std::vector<int> ins, outs_idx;
    Matrix fused;
    if (!rs_reconstruct_plan(enc->gen, k, present.data(), false, &ins, &outs_idx, &fused))
        return fail(SWEC_ERR_TOO_FEW_SHARDS, "not enough shards");
    const size_t chunk = std::max<size_t>(256, (size_t(std::min<int64_t>(int64_t(env_sz("SWEC_FILE_CHUNK", size_t(8) << 20)), std::max<int64_t>(todo, 1))) + 255) & ~size_t(255));
    FilePipeline pipe(enc, fused, chunk);
    if ((rc = pipe.start())) return rc;
    if (todo > 0 && worth_preallocating(out[size_t(outs_idx[0])]))
        pipe.parallel_for(int(outs_idx.size()), [&](int r) -> int {
            reserve_extents(out[size_t(outs_idx[size_t(r)])], todo);
            return 0;
        });
    for (int64_t o = 0; rc == SWEC_OK && o < todo; o += int64_t(chunk)) {
        Item it;
        it.len = size_t(std::min<int64_t>(int64_t(chunk), todo - o));
        for (int i = 0; i < k; i++) it.reads.push_back({i, in[size_t(ins[size_t(i)])], o, 0, it.len, in_d[size_t(ins[size_t(i)])]});
        for (size_t r = 0; r < outs_idx.size(); r++)
            it.writes.push_back({k + int(r), out[size_t(outs_idx[r])], o, out_d[size_t(outs_idx[r])]});
        rc = pipe.submit(std::move(it));
    }
    const int rc2 = pipe.finish();
    if (rc == SWEC_OK) rc = rc2;
    const std::string msg = pipe.error_message();
    pipe.shutdown();
    if (rc) {
        if (!msg.empty()) set_last_error(msg);
        return rc;
    }
    if (ragged) return fail(SWEC_ERR_SHARD_SIZE, "ec shard size expected 1048576 actual " + std::to_string(size % mib));
    undo.armed = false;
    return SWEC_OK;
}

int swec_verify_ec_files(const char* base, const char* const* dirs, int ndirs, int k, int m, int device,
                         uint64_t* mismatched_vectors, int* ok) {
    if (!base || !ok || (ndirs > 0 && !dirs)) return fail(SWEC_ERR_INVALID_ARG, "NULL argument");
    *ok = 0;
    const std::string b(base);
    if (k == 0) {
        int ds = 0, ps = 0;
        if (read_vif_ratio(b + ".vif", &ds, &ps) && ds > 0 && ps > 0 && ds + ps <= SWEC_MAX_SHARDS) { k = ds; m = ps; }
        else { k = 10; m = 4; }
    }
    swec_encoder* enc = nullptr;
    int rc = swec_encoder_new(k, m, device, &enc);
    if (rc) return rc;
    std::unique_ptr<swec_encoder, void (*)(swec_encoder*)> guard(enc, swec_encoder_free);
    const int total = k + m;
    std::string base_copy(b);
    const std::string base_name = basename(&base_copy[0]);
    FdSet fds;
    std::vector<int> in(static_cast<size_t>(total), -1);
    int64_t size = -1;
    for (int i = 0; i < total; i++) {  // verify needs every shard (verify_ec_shards, ec_encoder.rs:177-278)
        std::string path = b + shard_ext(i);
        if (!file_exists(path)) {
            path.clear();
            for (int d = 0; d < ndirs; d++) {
                const std::string cand = std::string(dirs[d]) + "/" + base_name + shard_ext(i);
                if (file_exists(cand)) { path = cand; break; }
            }
        }
        if (path.empty()) return fail(SWEC_ERR_TOO_FEW_SHARDS, "verify needs all shards; missing " + shard_ext(i));
        const int fd = open(path.c_str(), O_RDONLY);
        if (fd < 0) return io_fail("open " + path);
        fds.fds.push_back(fd);
        in[size_t(i)] = fd;
        struct stat st;
        if (fstat(fd, &st) != 0) return io_fail("fstat shard");
        if (size < 0) size = st.st_size;
        else if (size != st.st_size)
            return fail(SWEC_ERR_SHARD_SIZE, "ec shard size expected " + std::to_string(size) + " actual " + std::to_string(st.st_size));
    }
    Matrix rows(m, k);
    memcpy(rows.v.data(), enc->gen.row(k), rows.v.size());
    const size_t chunk = std::max<size_t>(256, (size_t(std::min<int64_t>(int64_t(env_sz("SWEC_FILE_CHUNK", size_t(8) << 20)), std::max<int64_t>(size, 1))) + 255) & ~size_t(255));
    FilePipeline pipe(enc, rows, chunk, /*verify=*/true);
    if ((rc = pipe.start())) return rc;
    for (int64_t o = 0; rc == SWEC_OK && o < size; o += int64_t(chunk)) {
        Item it;
        it.len = size_t(std::min<int64_t>(int64_t(chunk), size - o));
        for (int i = 0; i < total; i++) it.reads.push_back({i, in[size_t(i)], o, 0, it.len});
        rc = pipe.submit(std::move(it));
    }
    const int rc2 = pipe.finish();
    if (rc == SWEC_OK) rc = rc2;
    std::vector<unsigned long long> bad(static_cast<size_t>(m), 0);
    if (rc == SWEC_OK) rc = pipe.mismatches(bad.data());
    const std::string msg = pipe.error_message();
    pipe.shutdown();
    if (rc) {
        if (!msg.empty()) set_last_error(msg);
        return rc;
    }
    bool all_ok = true;
    for (int p = 0; p < m; p++) {
        if (mismatched_vectors) mismatched_vectors[p] = bad[size_t(p)];
        all_ok = all_ok && bad[size_t(p)] == 0;
    }
    *ok = all_ok ? 1 : 0;
    return SWEC_OK;
}

int swec_write_dat_file(const char* base, int64_t dat_size, const char* const* shard_names, int k, int64_t large,
                        int64_t small) {
    if (!base || !shard_names || k <= 0 || k > SWEC_MAX_SHARDS || large <= 0 || small <= 0 || dat_size < 0)
        return fail(SWEC_ERR_INVALID_ARG, "bad argument");
    FdSet fds;
    const int dat = open((std::string(base) + ".dat").c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
    if (dat < 0) return io_fail("cannot write volume .dat");
    fds.fds.push_back(dat);
    std::vector<int> in;
    for (int i = 0; i < k; i++) {
        const int fd = open(shard_names[i], O_RDONLY);
        if (fd < 0) return io_fail(std::string("open ") + shard_names[i]);
        fds.fds.push_back(fd);
        in.push_back(fd);
    }
    // The copy plan of the reference's two loops (ec_decoder.go:200-219) — shard s is read sequentially, the .dat is
    // written sequentially — as independent (shard offset → .dat offset) pieces, executed in parallel: every piece
    // has explicit offsets on both sides, so the order of execution cannot change the bytes.
    struct Piece { int shard; int64_t shard_off, dat_off, len; };
    std::vector<Piece> pieces;
    const int64_t max_piece = int64_t(8) << 20;
    std::vector<int64_t> pos(static_cast<size_t>(k), 0);
    int64_t out = 0, remaining = dat_size;
    auto plan = [&](int shard, int64_t n) {  // io.CopyN(datFile, inputFiles[shard], n)
        for (int64_t o = 0; o < n; o += max_piece)
            pieces.push_back({shard, pos[size_t(shard)] + o, out + o, std::min(max_piece, n - o)});
        pos[size_t(shard)] += n;
        out += n;
    };
    while (remaining >= int64_t(k) * large)
        for (int s = 0; s < k; s++) {
            plan(s, large);
            remaining -= large;
        }
    while (remaining > 0)
        for (int s = 0; s < k; s++) {
            const int64_t n = std::min(remaining, small);
            plan(s, n);
            remaining -= n;
        }
    // a shard shorter than the plan needs is the reference's "copy … block" error: check before writing anything
    for (int s2 = 0; s2 < k; s2++) {
        struct stat st;
        if (fstat(in[size_t(s2)], &st) != 0) return io_fail("fstat shard");
        if (st.st_size < pos[size_t(s2)]) return fail(SWEC_ERR_IO, "short read copying shard " + std::to_string(s2));
    }
    if (ftruncate(dat, off_t(dat_size)) != 0) return io_fail("size .dat");
    IoPool& pool = file_io_pool();
    std::mutex err_mu;
    std::string err_text;
    const std::function<int(int)> copy_piece = [&](int idx) -> int {
        const Piece& pc = pieces[size_t(idx)];
        int64_t done = 0;
        // kernel-side copy first (no user-space bounce; shares extents where the filesystem can) …
        while (done < pc.len) {
            off64_t oi = pc.shard_off + done, oo = pc.dat_off + done;
            const ssize_t n = copy_file_range(in[size_t(pc.shard)], &oi, dat, &oo, size_t(pc.len - done), 0);
            if (n <= 0) break;  // unsupported combination, or EOF: the read/write loop below decides
            done += n;
        }
        // … plain pread/pwrite for whatever is left
        std::vector<uint8_t> buf;
        while (done < pc.len) {
            if (buf.empty()) buf.resize(size_t(std::min<int64_t>(pc.len, int64_t(4) << 20)));
            const size_t want = size_t(std::min<int64_t>(pc.len - done, int64_t(buf.size())));
            const ssize_t got = pread(in[size_t(pc.shard)], buf.data(), want, off_t(pc.shard_off + done));
            if (got < 0 && errno == EINTR) continue;
            if (got <= 0) {
                std::lock_guard<std::mutex> lk(err_mu);
                if (err_text.empty()) err_text = "short read copying shard " + std::to_string(pc.shard);
                return SWEC_ERR_IO;
            }
            ssize_t put = 0;
            while (put < got) {
                const ssize_t w = pwrite(dat, buf.data() + put, size_t(got - put), off_t(pc.dat_off + done + put));
                if (w < 0 && errno == EINTR) continue;
                if (w <= 0) {
                    std::lock_guard<std::mutex> lk(err_mu);
                    if (err_text.empty()) err_text = std::string("write .dat: ") + strerror(errno);
                    return SWEC_ERR_IO;
                }
                put += w;
            }
            done += got;
        }
        return SWEC_OK;
    };
    const int rc = pool.parallel_for(int(pieces.size()), copy_piece);
    if (rc) return fail(rc, err_text);
    return SWEC_OK;
}

// ---- layout arithmetic -----------------------------------------------------------------------

int swec_locate_data(int64_t large, int64_t small, int64_t shard_dat_size, int64_t offset, int64_t size, int k,
                     swec_interval* out, int cap) {
    if (large <= 0 || small <= 0 || k <= 0 || !out) return fail(SWEC_ERR_INVALID_ARG, "bad argument");
    const int64_t nlarge_rows = shard_dat_size / large;  // ec_locate.go:67
    const int64_t large_area = nlarge_rows * large * k;
    bool is_large = offset < large_area;
    const int64_t rel = is_large ? offset : offset - large_area;
    const int64_t blk = is_large ? large : small;
    int64_t block_index = rel / blk, inner = rel % blk;
    int n = 0;
    while (size > 0) {
        const int64_t room = (is_large ? large : small) - inner;
        if (room > 0) {
            if (n >= cap) return fail(SWEC_ERR_INVALID_ARG, "interval buffer too small");
            swec_interval& iv = out[n++];
            iv.block_index = int32_t(block_index);
            iv.is_large_block = is_large ? 1 : 0;
            iv.inner_block_offset = inner;
            iv.large_block_rows_count = int32_t(nlarge_rows);
            iv.reserved = 0;
            iv.size = std::min(size, room);
            size -= iv.size;
            if (size == 0) break;
        }
        // moveToNextBlock (ec_locate.go:55-63): the block after the last large one is small block 0
        block_index++;
        if (is_large && block_index == nlarge_rows * k) {
            is_large = false;
            block_index = 0;
        }
        inner = 0;
    }
    return n;
}

void swec_interval_to_shard(const swec_interval* iv, int64_t large, int64_t small, int k, int* shard_id,
                            int64_t* shard_offset) {
    const int64_t row = iv->block_index / k;  // ec_locate.go:87-98
    int64_t off = iv->inner_block_offset;
    off += iv->is_large_block ? row * large : int64_t(iv->large_block_rows_count) * large + row * small;
    if (shard_id) *shard_id = iv->block_index % k;
    if (shard_offset) *shard_offset = off;
}

}  // extern "C"
