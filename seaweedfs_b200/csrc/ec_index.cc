// seaweedfs_b200/csrc/ec_index.cc — the index files either side of the RS path (SURVEY §8f row 4,
// Appendix C).  No GF arithmetic here: host-only twins of
//   WriteSortedFileFromIdx / readNeedleMap        weed/storage/erasure_coding/ec_encoder.go:31-58,379-396
//   RebuildEcxFile / MarkNeedleDeleted            weed/storage/erasure_coding/ec_volume_delete.go:13-26,95-142
//   SearchNeedleFromSortedIndex                   weed/storage/erasure_coding/ec_volume.go:431-458
//   WriteIdxFileFromEcIndex, FindDatFileSize, HasLiveNeedles   weed/storage/erasure_coding/ec_decoder.go:23-92
// so that a volume server using libswec for ec.encode / ec.rebuild / ec.decode needs nothing else
// from the Go package for these files.  Entry format (4-byte offsets, the default build):
// 8-byte needle id, 4-byte offset in units of 8 bytes, 4-byte size, all big-endian
// (weed/storage/types/needle_types.go:58-64, offset_4bytes.go:14-60, needle_map/needle_value.go:24-30).
#include <errno.h>
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "engine.h"

namespace swec {
namespace {

constexpr int kEntry = 16;             // NeedleMapEntrySize
constexpr int32_t kTombstone = -1;     // TombstoneFileSize
constexpr int64_t kSuperBlockSize = 8; // super_block.SuperBlockSize

uint64_t be64(const uint8_t* p) {
    uint64_t v = 0;
    for (int i = 0; i < 8; i++) v = (v << 8) | p[i];
    return v;
}
uint32_t be32(const uint8_t* p) { return (uint32_t(p[0]) << 24) | (uint32_t(p[1]) << 16) | (uint32_t(p[2]) << 8) | p[3]; }
void put_be64(uint8_t* p, uint64_t v) {
    for (int i = 7; i >= 0; i--) { p[i] = uint8_t(v); v >>= 8; }
}
void put_be32(uint8_t* p, uint32_t v) {
    for (int i = 3; i >= 0; i--) { p[i] = uint8_t(v); v >>= 8; }
}
bool size_deleted(int32_t s) { return s < 0 || s == kTombstone; }  // Size.IsDeleted, needle_types.go:25-27

bool read_all(const std::string& path, std::vector<uint8_t>* out) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    uint8_t buf[1 << 16];
    size_t n;
    out->clear();
    while ((n = fread(buf, 1, sizeof buf, f)) > 0) out->insert(out->end(), buf, buf + n);
    fclose(f);
    return true;
}

// Replaces `path` atomically: the bytes go to <path>.tmp.<pid>, which is renamed over the target.  A mounted
// EcVolume maps .ecx MAP_SHARED (ec_volume.cc) and binary-searches the mapping; rewriting the same inode in place
// (O_TRUNC) while it is mounted — ec.encode re-run on a mounted volume — would leave the search touching pages past
// the new EOF: SIGBUS, which kills the whole volume server where the reference's ReadAt only returns an error.
// With rename() existing mappings keep the old inode and its bytes until they are unmapped.
bool write_all(const std::string& path, const std::vector<uint8_t>& data) {
    static std::atomic<unsigned> serial{0};  // two threads of one process may rewrite the same path
    const std::string tmp = path + ".tmp." + std::to_string(long(getpid())) + "." + std::to_string(serial++);
    const int fd = open(tmp.c_str(), O_TRUNC | O_CREAT | O_WRONLY, 0644);
    if (fd < 0) return false;
    size_t put = 0;
    while (put < data.size()) {
        const ssize_t n = write(fd, data.data() + put, data.size() - put);
        if (n < 0) {
            if (errno == EINTR) continue;
            const int keep = errno;
            close(fd);
            unlink(tmp.c_str());
            errno = keep;
            return false;
        }
        put += size_t(n);
    }
    if (close(fd) != 0 || rename(tmp.c_str(), path.c_str()) != 0) {
        const int keep = errno;
        unlink(tmp.c_str());
        errno = keep;
        return false;
    }
    return true;
}

bool exists(const std::string& p) {
    struct stat st;
    return stat(p.c_str(), &st) == 0;
}

int io_err(const std::string& what) { return fail(SWEC_ERR_IO, what + ": " + strerror(errno)); }

// GetActualSize (needle/needle_read.go:292-294, needle_read_tail.go:36-50): header 16 + body + checksum 4
// (+ 8-byte timestamp in version 3) + padding to 8, where the padding is 1..8 bytes, never 0.
int64_t actual_size(int32_t size, int version) {
    const int64_t fixed = 16 + int64_t(size) + 4 + (version == 3 ? 8 : 0);
    return fixed + (8 - fixed % 8);
}

}  // namespace
}  // namespace swec

using namespace swec;

extern "C" {

int swec_write_sorted_file_from_idx(const char* base, const char* ext) {
    if (!base || !ext) return fail(SWEC_ERR_INVALID_ARG, "NULL argument");
    std::vector<uint8_t> idx;
    if (!read_all(std::string(base) + ".idx", &idx)) return io_err(std::string("cannot read Volume Index ") + base + ".idx");
    // readNeedleMap replays the .idx into a map: a live entry (non-zero offset, size not deleted) sets the key, anything
    // else removes it — so the LAST entry of a key alone decides whether and how the key appears.  Sorting the entries
    // by (key, position) and keeping each key's last one gives the same file without a 30-million-node tree for a
    // full 30 GB volume of small needles.
    struct E { uint64_t key; uint32_t pos, offset, size; };
    const size_t n = idx.size() / kEntry;
    if (n > 0xFFFFFFFFull) return fail(SWEC_ERR_INVALID_ARG, "index too large");
    std::vector<E> es(n);
    for (size_t i = 0; i < n; i++) {
        const uint8_t* p = &idx[i * kEntry];
        es[i] = {be64(p), uint32_t(i), be32(p + 8), be32(p + 12)};
    }
    std::sort(es.begin(), es.end(), [](const E& a, const E& b) { return a.key != b.key ? a.key < b.key : a.pos < b.pos; });
    std::vector<uint8_t> out;
    out.reserve(n * kEntry);
    for (size_t i = 0; i < n; i++) {
        if (i + 1 < n && es[i + 1].key == es[i].key) continue;  // not the key's last word
        const E& e = es[i];
        if (e.offset == 0 || size_deleted(int32_t(e.size))) continue;  // the key ends deleted
        uint8_t rec[kEntry];
        put_be64(rec, e.key);
        put_be32(rec + 8, e.offset);
        put_be32(rec + 12, e.size);
        out.insert(out.end(), rec, rec + kEntry);  // ascending keys: AscendingVisit
    }
    if (!write_all(std::string(base) + ext, out)) return io_err("failed to open ecx file");
    return SWEC_OK;
}

int swec_rebuild_ecx_file(const char* base) {
    if (!base) return fail(SWEC_ERR_INVALID_ARG, "NULL argument");
    const std::string b(base);
    if (!exists(b + ".ecj")) return SWEC_OK;
    const int ecx = open((b + ".ecx").c_str(), O_RDWR);
    if (ecx < 0) return io_err("rebuild: failed to open ecx file");
    std::vector<uint8_t> index, ecj;
    if (!read_all(b + ".ecx", &index)) {
        close(ecx);
        return io_err("rebuild: failed to read ecx file");
    }
    const int64_t entries = int64_t(index.size()) / kEntry;
    if (!read_all(b + ".ecj", &ecj)) {
        close(ecx);
        return io_err("rebuild: failed to open ecj file");
    }
    // SearchNeedleFromSortedIndex + MarkNeedleDeleted for every journalled id: the search runs on the copy in
    // memory (a 30 GB volume of small needles has a 480 MB index and 25 probes per id), the tombstone is written
    // in place on disk, exactly the four size bytes the reference rewrites
    for (size_t off = 0; off + 8 <= ecj.size(); off += 8) {
        const uint64_t id = be64(&ecj[off]);
        int64_t lo = 0, hi = entries;
        while (lo < hi) {
            const int64_t mid = (lo + hi) / 2;
            const uint64_t key = be64(&index[size_t(mid) * kEntry]);
            if (key == id) {
                uint8_t t[4];
                put_be32(t, uint32_t(kTombstone));
                if (memcmp(&index[size_t(mid) * kEntry + 12], t, 4) != 0) {
                    memcpy(&index[size_t(mid) * kEntry + 12], t, 4);
                    if (pwrite(ecx, t, 4, off_t(mid * kEntry + 12)) != 4) {
                        close(ecx);
                        return io_err("sorted needle write error");
                    }
                }
                break;
            }
            if (key < id) lo = mid + 1;
            else hi = mid;
        }
    }
    close(ecx);
    unlink((b + ".ecj").c_str());
    return SWEC_OK;
}

int swec_write_idx_file_from_ec_index(const char* base) {
    if (!base) return fail(SWEC_ERR_INVALID_ARG, "NULL argument");
    const std::string b(base);
    std::vector<uint8_t> data;
    if (!read_all(b + ".ecx", &data)) return io_err("cannot open ec index " + b + ".ecx");
    std::vector<uint8_t> ecj;
    if (exists(b + ".ecj") && !read_all(b + ".ecj", &ecj)) return io_err("cannot open ec index " + b + ".ecj");
    for (size_t off = 0; off + 8 <= ecj.size(); off += 8) {  // one tombstone entry per journalled id
        uint8_t e[kEntry] = {0};
        memcpy(e, &ecj[off], 8);
        put_be32(e + 12, uint32_t(kTombstone));
        data.insert(data.end(), e, e + kEntry);
    }
    if (!write_all(b + ".idx", data)) return io_err("cannot open " + b + ".idx");
    return SWEC_OK;
}

int swec_has_live_needles(const char* index_base, int* has_live) {
    if (!index_base || !has_live) return fail(SWEC_ERR_INVALID_ARG, "NULL argument");
    std::vector<uint8_t> ecx;
    if (!read_all(std::string(index_base) + ".ecx", &ecx)) return io_err(std::string("cannot open ec index ") + index_base + ".ecx");
    *has_live = 0;
    for (size_t off = 0; off + kEntry <= ecx.size(); off += kEntry)
        if (!size_deleted(int32_t(be32(&ecx[off + 12])))) {
            *has_live = 1;
            break;
        }
    return SWEC_OK;
}

// idx.CheckIndexFile (weed/storage/idx/check.go:36-111) — what EcVolume.ScrubIndex runs on .ecx
// (ec_volume_scrub.go:20-25): entries sorted by (offset, size); two neighbours overlap when the later one starts at
// or before the end of the earlier one; and the file must be a whole number of entries.
int swec_check_index_file(const char* path, int needle_version, int64_t* entries, char* errors, size_t errors_cap,
                          int* n_errors) {
    if (!path || !entries || !n_errors) return fail(SWEC_ERR_INVALID_ARG, "NULL argument");
    std::vector<uint8_t> raw;
    if (!read_all(path, &raw)) return io_err(std::string("cannot read index ") + path);
    struct E { int index; uint64_t id; int64_t offset; int32_t size; };
    std::vector<E> es;
    for (size_t off = 0; off + kEntry <= raw.size(); off += kEntry)  // WalkIndexFile ignores a trailing partial entry
        es.push_back({int(es.size()), be64(&raw[off]), int64_t(be32(&raw[off + 8])) * 8, int32_t(be32(&raw[off + 12]))});
    std::stable_sort(es.begin(), es.end(), [](const E& a, const E& b) { return a.offset != b.offset ? a.offset < b.offset : a.size < b.size; });
    // needle.GetActualSize with the reference's types: PaddingLength is computed in Size (int32) arithmetic and wraps
    // like Go does; NeedleBodyLength adds in int64 (needle_read_tail.go:36-49)
    auto actual = [&](int32_t size) -> int64_t {
        const uint32_t tail = needle_version == 3 ? 4u + 8u : 4u;
        const int32_t sum = int32_t(16u + uint32_t(size) + tail);  // wraps
        const int32_t padding = 8 - (sum % 8);                      // Go's % keeps the sign of the dividend, like C++
        return 16 + int64_t(size) + int64_t(tail) + int64_t(padding);
    };
    std::string text;
    int count = 0;
    auto add = [&](const std::string& m) {
        if (count++) text += "\n";
        text += m;
    };
    for (size_t i = 1; i < es.size(); i++) {
        const E &e = es[i], &last = es[i - 1];
        int64_t end = e.offset, last_end = last.offset;
        if (const int64_t sz = actual(e.size)) end += sz - 1;
        if (const int64_t sz = actual(last.size)) last_end += sz - 1;
        if (e.offset <= last_end)
            add("needle " + std::to_string(e.id) + " (#" + std::to_string(e.index + 1) + ") at [" + std::to_string(e.offset) + "-" +
                std::to_string(end) + "] overlaps needle " + std::to_string(last.id) + " at [" + std::to_string(last.offset) + "-" +
                std::to_string(last_end) + "]");
    }
    const int64_t n = int64_t(es.size());
    if (n * kEntry != int64_t(raw.size()))
        add("expected an index file of size " + std::to_string(raw.size()) + ", got " + std::to_string(n * kEntry));
    *entries = n;
    *n_errors = count;
    if (errors && errors_cap) {
        const size_t m = std::min(text.size(), errors_cap - 1);
        memcpy(errors, text.data(), m);
        errors[m] = 0;
    }
    return SWEC_OK;
}

int swec_find_dat_file_size(const char* data_base, const char* index_base, int64_t* dat_size) {
    if (!data_base || !index_base || !dat_size) return fail(SWEC_ERR_INVALID_ARG, "NULL argument");
    // readEcVolumeVersion: the superblock sits at the start of .ec00; byte 0 is the needle version
    const int fd = open((std::string(data_base) + ".ec00").c_str(), O_RDONLY);
    if (fd < 0) return io_err(std::string("open ec volume ") + data_base + " superblock");
    uint8_t sb[kSuperBlockSize];
    const ssize_t got = pread(fd, sb, sizeof sb, 0);
    close(fd);
    if (got != ssize_t(sizeof sb)) return fail(SWEC_ERR_IO, "cannot read the superblock from .ec00");
    const int version = sb[0];
    std::vector<uint8_t> ecx;
    if (!read_all(std::string(index_base) + ".ecx", &ecx)) return io_err(std::string("cannot open ec index ") + index_base + ".ecx");
    int64_t size = kSuperBlockSize;
    for (size_t off = 0; off + kEntry <= ecx.size(); off += kEntry) {
        const int32_t sz = int32_t(be32(&ecx[off + 12]));
        if (size_deleted(sz)) continue;
        const int64_t stop = int64_t(be32(&ecx[off + 8])) * 8 + actual_size(sz, version);
        if (stop > size) size = stop;
    }
    *dat_size = size;
    return SWEC_OK;
}

}  // extern "C"
