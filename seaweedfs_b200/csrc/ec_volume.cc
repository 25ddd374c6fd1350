// seaweedfs_b200/csrc/ec_volume.cc — the volume-server-local bodies of the three gRPC handlers that
// drive the RS path, as single C-ABI calls:
//   swec_ec_shards_generate    VolumeEcShardsGenerate   weed/server/volume_grpc_erasure_coding.go:43-146
//   swec_ec_shards_rebuild     VolumeEcShardsRebuild    weed/server/volume_grpc_erasure_coding.go:149-225
//   swec_ec_shards_to_volume   VolumeEcShardsToVolume   weed/server/volume_grpc_erasure_coding.go:578-668
//   swec_read_ec_needles       Store.ReadEcShardNeedle + readEcShardIntervals + readOneEcShardInterval +
//                              recoverOneRemoteEcShardInterval, on local shard files, batched
//                                                       weed/storage/store_ec.go:252-355,482-560
// Everything the handlers do to FILES is here, in the reference's order (.ecx before the shards, the
// .dat size snapshot before encoding, .vif last, partial outputs removed on any error); what they do
// to the server's in-memory state (volume lookup, maintenance mode, disk-location scan, compaction)
// stays in Go.  The shard arithmetic runs on the GPU through swec_generate_ec_files /
// swec_rebuild_ec_files; the index and .vif work is host-only.
#include <errno.h>
#include <fcntl.h>
#include <libgen.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cctype>
#include <cinttypes>
#include <cstdio>
#include <cstring>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "engine.h"
#include "mini_json.h"

namespace swec {

namespace {

std::string ext_of(int idx) {  // ToExt, ec_encoder.go:106-108
    char b[16];
    snprintf(b, sizeof b, ".ec%02d", idx);
    return b;
}

bool is_file(const std::string& p) {
    struct stat st;
    return stat(p.c_str(), &st) == 0 && !S_ISDIR(st.st_mode);
}

// SaveVolumeInfo (weed/storage/volume_info/volume_info.go:73-95): protojson with EmitUnpopulated and
// a two-space indent.  protojson renders 64-bit integers as strings and deliberately does not promise
// byte-stable whitespace, so readers (ours: read_vif_ratio; Go: protojson.Unmarshal) parse, not compare.
int save_volume_info(const std::string& path, uint32_t version, int64_t dat_size, uint64_t expire_at_sec, int ds,
                     int ps) {
    struct stat st;
    if (stat(path.c_str(), &st) == 0 && access(path.c_str(), W_OK) != 0)
        return fail(SWEC_ERR_IO, "failed to check " + path + " not writable");
    char text[512];
    const int n = snprintf(text, sizeof text,
                           "{\n"
                           "  \"files\": [],\n"
                           "  \"version\": %u,\n"
                           "  \"replication\": \"\",\n"
                           "  \"bytesOffset\": 0,\n"
                           "  \"datFileSize\": \"%" PRId64 "\",\n"
                           "  \"expireAtSec\": \"%" PRIu64 "\",\n"
                           "  \"readOnly\": false,\n"
                           "  \"ecShardConfig\": {\n"
                           "    \"dataShards\": %d,\n"
                           "    \"parityShards\": %d\n"
                           "  }\n"
                           "}",
                           version, dat_size, expire_at_sec, ds, ps);
    const int fd = open(path.c_str(), O_TRUNC | O_CREAT | O_WRONLY, 0644);
    if (fd < 0) return fail(SWEC_ERR_IO, "failed to write " + path + ": " + strerror(errno));
    int put = 0;
    while (put < n) {
        const ssize_t w = write(fd, text + put, size_t(n - put));
        if (w < 0) {
            if (errno == EINTR) continue;
            const int e = errno;
            close(fd);
            return fail(SWEC_ERR_IO, "failed to write " + path + ": " + strerror(e));
        }
        put += int(w);
    }
    close(fd);
    return SWEC_OK;
}

// the EC ratio a handler works with: an existing valid .vif wins, else 10+4
// (volume_grpc_erasure_coding.go:61-77, ec_encoder.go:76-95, ec_volume.go:114-154)
void ratio_from_vif(const std::string& data_base, int* k, int* m) {
    int ds = 0, ps = 0;
    if (read_vif_ratio(data_base + ".vif", &ds, &ps) && ds > 0 && ps > 0 && ds + ps <= SWEC_MAX_SHARDS) {
        *k = ds;
        *m = ps;
    } else {
        *k = 10;
        *m = 4;
    }
}

// A numeric field of a protobuf-JSON .vif (64-bit integers are rendered as strings); false when absent.
bool vif_number(const std::string& txt, const char* key, int64_t* out) {
    const std::string k(key);
    const char* alt = k == "datFileSize" ? "dat_file_size" : k == "expireAtSec" ? "expire_at_sec" : nullptr;
    return mini_json::top_int(txt, key, alt, out);
}

bool slurp(const std::string& path, std::string* out) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    char buf[1 << 16];
    size_t n;
    out->clear();
    while ((n = fread(buf, 1, sizeof buf, f)) > 0) out->append(buf, n);
    fclose(f);
    return true;
}

uint64_t be64(const uint8_t* p) {
    uint64_t v = 0;
    for (int i = 0; i < 8; i++) v = (v << 8) | p[i];
    return v;
}
uint32_t be32(const uint8_t* p) { return (uint32_t(p[0]) << 24) | (uint32_t(p[1]) << 16) | (uint32_t(p[2]) << 8) | p[3]; }

// GetActualSize (needle/needle_read.go:292-294, needle_read_tail.go:36-49)
int64_t needle_actual_size(int64_t size, int version) {
    const int64_t fixed = 16 + size + 4 + (version == 3 ? 8 : 0);
    return fixed + (8 - fixed % 8);
}

}  // namespace
}  // namespace swec

using namespace swec;

extern "C" {

int swec_ec_shards_generate(const char* data_base, const char* index_base, uint32_t needle_version,
                            uint64_t expire_at_sec, int device) {
    if (!data_base) return fail(SWEC_ERR_INVALID_ARG, "data_base_file_name is NULL");
    const std::string db(data_base), ib(index_base && *index_base ? index_base : data_base);
    int k, m;
    ratio_from_vif(db, &k, &m);

    struct Cleanup {  // the handler's deferred cleanup: shards and .ecx go away unless we reach the end
        const std::string &db, &ib;
        int total;
        bool armed = true;
        ~Cleanup() {
            if (!armed) return;
            const std::string keep = last_error();  // unlink() must not disturb the reported detail
            for (int i = 0; i < total; i++) unlink((db + ext_of(i)).c_str());
            unlink((ib + ".ecx").c_str());
            set_last_error(keep);
        }
    } cleanup{db, ib, k + m};

    // .ecx BEFORE the shards (the race the reference documents at :82-95)
    int rc = swec_write_sorted_file_from_idx(ib.c_str(), ".ecx");
    if (rc) return rc;
    // snapshot of the .dat size before encoding — what .ecx references (:103)
    struct stat st;
    if (stat((db + ".dat").c_str(), &st) != 0) return fail(SWEC_ERR_IO, "failed to stat dat file " + db + ".dat: " + strerror(errno));
    if (needle_version == 0) {  // v.Version(): byte 0 of the superblock (super_block.go; ec_decoder.go:94-111)
        const int fd = open((db + ".dat").c_str(), O_RDONLY);
        uint8_t b0 = 0;
        if (fd < 0 || pread(fd, &b0, 1, 0) != 1) {
            const int e = errno;
            if (fd >= 0) close(fd);
            return fail(SWEC_ERR_IO, "cannot read the superblock of " + db + ".dat: " + strerror(e));
        }
        close(fd);
        needle_version = b0;
    }
    // WriteEcFilesWithContext: 256 KiB buffers, 1 GiB / 1 MiB blocks (ec_encoder.go:67-69)
    rc = swec_generate_ec_files(db.c_str(), 256 * 1024, int64_t(1) << 30, int64_t(1) << 20, k, m, device);
    if (rc) return rc;
    rc = save_volume_info(db + ".vif", needle_version, int64_t(st.st_size), expire_at_sec, k, m);
    if (rc) return rc;
    cleanup.armed = false;
    return SWEC_OK;
}

int swec_ec_shards_rebuild(const char* data_base, const char* index_base, const char* const* additional_dirs,
                           int n_additional_dirs, int device, uint32_t* rebuilt, int* n_rebuilt) {
    if (!data_base || !rebuilt || !n_rebuilt) return fail(SWEC_ERR_INVALID_ARG, "NULL argument");
    // RebuildEcFiles: ratio from data_base.vif or 10+4; inputs searched in data_base's directory, then
    // additional_dirs; outputs created next to data_base (:203-209)
    int rc = swec_rebuild_ec_files(data_base, additional_dirs, n_additional_dirs, 0, 0, device, rebuilt, n_rebuilt);
    if (rc) return rc;
    // RebuildEcxFile on the index base, falling back to the data directory (:211-217)
    std::string ib(index_base && *index_base ? index_base : data_base);
    if (!is_file(ib + ".ecx") && ib != data_base) ib = data_base;
    return swec_rebuild_ecx_file(ib.c_str());
}

int swec_ec_shards_to_volume(const char* data_base, const char* index_base, const char* const* additional_dirs,
                             int n_additional_dirs, int64_t* dat_file_size) {
    if (!data_base || (n_additional_dirs > 0 && !additional_dirs)) return fail(SWEC_ERR_INVALID_ARG, "NULL argument");
    const std::string db(data_base);
    int k, m;
    ratio_from_vif(db, &k, &m);  // NewEcVolume loads the ratio from .vif (ec_volume.go:114-154)
    // CollectEcShards: every data shard must be found locally (:601-606)
    std::string base_copy(db);
    const std::string base_name = basename(&base_copy[0]);
    std::vector<std::string> names;
    for (int i = 0; i < k; i++) {
        std::string path = db + ext_of(i);
        if (!is_file(path)) {
            path.clear();
            for (int d = 0; d < n_additional_dirs; d++) {
                const std::string cand = std::string(additional_dirs[d]) + "/" + base_name + ext_of(i);
                if (is_file(cand)) {
                    path = cand;
                    break;
                }
            }
        }
        if (path.empty()) return fail(SWEC_ERR_TOO_FEW_SHARDS, "ec volume missing shard " + std::to_string(i));
        names.push_back(path);
    }
    std::string ib(index_base && *index_base ? index_base : data_base);
    if (!is_file(ib + ".ecx")) ib = db;  // :608-611
    int rc = swec_rebuild_ecx_file(ib.c_str());  // fold .ecj first so deleted needles are not counted live
    if (rc) return rc;
    int live = 0;
    if ((rc = swec_has_live_needles(ib.c_str(), &live))) return rc;
    if (!live) return fail(SWEC_ERR_NO_LIVE_NEEDLES, "ec volume has no live entries");  // EcNoLiveEntriesSubstring
    int64_t size = 0;
    // FindDatFileSize reads the needle version from <data_base>.ec00 (ec_decoder.go:94-111)
    std::string ec00_base = names[0].substr(0, names[0].size() - 5);
    if ((rc = swec_find_dat_file_size(ec00_base.c_str(), ib.c_str(), &size))) return rc;
    std::vector<const char*> cnames;
    for (const auto& s : names) cnames.push_back(s.c_str());
    if ((rc = swec_write_dat_file(db.c_str(), size, cnames.data(), k, int64_t(1) << 30, int64_t(1) << 20))) return rc;
    if ((rc = swec_write_idx_file_from_ec_index(ib.c_str()))) return rc;
    if (dat_file_size) *dat_file_size = size;
    return SWEC_OK;
}

// ---- EcVolume: the mounted state the read path works from (ec_volume.go:36-160) ------------------

}  // extern "C"

struct swec_ec_volume {
    std::mutex mu;
    int k = 10, m = 4, version = 3, device = 0;
    int64_t shard_dat_size = 0;
    std::string index_base;
    std::vector<int> shard_fd;  // total entries, -1 = not local
    // the sealed index is mapped, not copied: a full volume of small needles has a ~500 MB .ecx and a server mounts
    // hundreds of EC volumes; a lookup touches ~25 pages of the page cache (the reference does 25 ReadAt calls)
    const uint8_t* ecx_map = nullptr;
    size_t ecx_bytes = 0;
    std::string ecj;
    std::vector<uint64_t> deleted;  // ids of .ecj, sorted and unique: the reference's in-memory deletedNeedles set
    int64_t ecj_size_seen = -1, ecj_mtime_ns_seen = 0;
    uint64_t ecj_inode_seen = 0;
    void index_journal() {
        deleted.clear();
        for (size_t off = 0; off + 8 <= ecj.size(); off += 8) {
            uint64_t v = 0;
            for (int i = 0; i < 8; i++) v = (v << 8) | uint8_t(ecj[off + size_t(i)]);
            deleted.push_back(v);
        }
        std::sort(deleted.begin(), deleted.end());
        deleted.erase(std::unique(deleted.begin(), deleted.end()), deleted.end());
    }
    void refresh_journal(bool force = false);
    swec_encoder* enc = nullptr;  // created on the first recovery, keeps its staging ring and kernels
    ~swec_ec_volume() {
        if (ecx_map && ecx_bytes) munmap(const_cast<uint8_t*>(ecx_map), ecx_bytes);
        for (int fd : shard_fd)
            if (fd >= 0) close(fd);
        if (enc) swec_encoder_free(enc);
    }
};

void swec_ec_volume::refresh_journal(bool force) {
    // the deletion journal grows while the volume is mounted (DeleteNeedleFromEcx appends to .ecj): pick up new
    // entries when the file moved — size, mtime or inode, so a journal folded and re-created to the same length is
    // seen too — the reference keeps the same set in memory (ec_volume.go:351-384)
    struct stat st;
    const bool have = stat((index_base + ".ecj").c_str(), &st) == 0;
    const int64_t now = have ? int64_t(st.st_size) : 0;
    const int64_t mt = have ? int64_t(st.st_mtim.tv_sec) * 1000000000ll + st.st_mtim.tv_nsec : 0;
    const uint64_t ino = have ? uint64_t(st.st_ino) : 0;
    if (!force && now == ecj_size_seen && mt == ecj_mtime_ns_seen && ino == ecj_inode_seen) return;
    if (now == 0 || !swec::slurp(index_base + ".ecj", &ecj)) ecj.clear();
    ecj_size_seen = now;
    ecj_mtime_ns_seen = mt;
    ecj_inode_seen = ino;
    index_journal();
}

extern "C" {

int swec_ec_volume_open(const char* data_base, const char* index_base, const char* const* additional_dirs,
                        int n_additional_dirs, int device, swec_ec_volume** out) {
    if (!out) return fail(SWEC_ERR_INVALID_ARG, "out is NULL");
    *out = nullptr;
    if (!data_base || (n_additional_dirs > 0 && !additional_dirs)) return fail(SWEC_ERR_INVALID_ARG, "NULL argument");
    const std::string db(data_base);
    std::string ib(index_base && *index_base ? index_base : data_base);
    if (!is_file(ib + ".ecx")) ib = db;  // NewEcVolume falls back to the data directory (ec_volume.go:72-85)
    std::unique_ptr<swec_ec_volume> v(new (std::nothrow) swec_ec_volume());
    if (!v) return fail(SWEC_ERR_NOMEM, "out of memory");
    v->device = device;
    v->index_base = ib;

    // what NewEcVolume loads: ratio, needle version and datFileSize from .vif (ec_volume.go:114-154)
    ratio_from_vif(db, &v->k, &v->m);
    const int total = v->k + v->m;
    int64_t dat_file_size = 0;
    {
        std::string vif;
        if (slurp(db + ".vif", &vif) || slurp(ib + ".vif", &vif)) {
            int64_t x = 0;
            if (vif_number(vif, "version", &x) && x > 0) v->version = int(x);
            if (vif_number(vif, "datFileSize", &x)) dat_file_size = x;
        }
    }
    // local shards: data_base's directory, then the other disks
    std::string base_copy(db);
    const std::string base_name = basename(&base_copy[0]);
    v->shard_fd.assign(size_t(total), -1);
    int64_t ecd_file_size = -1;
    int nlocal = 0;
    for (int i = 0; i < total; i++) {
        std::string path = db + ext_of(i);
        if (!is_file(path)) {
            path.clear();
            for (int d = 0; d < n_additional_dirs; d++) {
                const std::string cand = std::string(additional_dirs[d]) + "/" + base_name + ext_of(i);
                if (is_file(cand)) {
                    path = cand;
                    break;
                }
            }
        }
        if (path.empty()) continue;
        const int fd = open(path.c_str(), O_RDONLY);
        if (fd < 0) continue;  // unreadable = not local; its intervals are recovered from the others
        v->shard_fd[size_t(i)] = fd;
        nlocal++;
        if (ecd_file_size < 0) {
            struct stat st;
            if (fstat(fd, &st) == 0) ecd_file_size = st.st_size;
        }
    }
    if (nlocal == 0) return fail(SWEC_ERR_TOO_FEW_SHARDS, "ec shard " + db + " not found");
    // LocateEcShardNeedleInterval: .vif's datFileSize is authoritative; old volumes fall back to the
    // shard file size minus one (ec_volume.go:399-417)
    v->shard_dat_size = dat_file_size > 0 ? dat_file_size / v->k : ecd_file_size - 1;
    {
        const int efd = open((ib + ".ecx").c_str(), O_RDONLY);
        if (efd < 0) return fail(SWEC_ERR_IO, "cannot open ec volume index " + ib + ".ecx: " + strerror(errno));
        struct stat st;
        if (fstat(efd, &st) != 0) {
            const int e = errno;
            close(efd);
            return fail(SWEC_ERR_IO, "can not stat ec volume index " + ib + ".ecx: " + strerror(e));
        }
        v->ecx_bytes = size_t(st.st_size);
        if (v->ecx_bytes) {  // MAP_SHARED: tombstones that RebuildEcxFile writes in place are seen
            void* m = mmap(nullptr, v->ecx_bytes, PROT_READ, MAP_SHARED, efd, 0);
            if (m == MAP_FAILED) {
                const int e = errno;
                close(efd);
                v->ecx_bytes = 0;
                return fail(SWEC_ERR_IO, "cannot map ec volume index " + ib + ".ecx: " + strerror(e));
            }
            v->ecx_map = static_cast<const uint8_t*>(m);
        }
        close(efd);
    }
    *out = v.release();
    return SWEC_OK;
}

void swec_ec_volume_close(swec_ec_volume* v) { delete v; }

int swec_ec_volume_read_needles(swec_ec_volume* v, swec_needle_read* reads, int n_reads) {
    if (!v || (n_reads > 0 && !reads)) return fail(SWEC_ERR_INVALID_ARG, "NULL argument");
    std::lock_guard<std::mutex> lock(v->mu);
    const int k = v->k, total = v->k + v->m, version = v->version;
    const int64_t large = int64_t(1) << 30, small = int64_t(1) << 20;

    v->refresh_journal();
    const uint8_t* ex = v->ecx_map;
    const int64_t entries = int64_t(v->ecx_bytes) / 16;
    auto journalled = [&](uint64_t id) { return std::binary_search(v->deleted.begin(), v->deleted.end(), id); };

    // ---- pass 1: locate every needle, read what is local, collect what must be recovered
    struct Recover {
        int read_idx;
        size_t buf_off, len;
        int shard;
        std::vector<std::vector<uint8_t>> bufs;  // total entries; empty = not available
        std::vector<uint8_t*> ptrs;
        std::vector<uint8_t> present;
    };
    std::vector<Recover> recs;
    for (int r = 0; r < n_reads; r++) {
        swec_needle_read& rd = reads[r];
        rd.offset = 0;
        rd.size = 0;
        rd.n_bytes = 0;
        rd.n_recovered_intervals = 0;
        rd.status = SWEC_OK;
        int64_t lo = 0, hi = entries, found = -1;  // SearchNeedleFromSortedIndex (ec_volume.go:431-458)
        while (lo < hi) {
            const int64_t mid = (lo + hi) / 2;
            const uint64_t key = be64(ex + mid * 16);
            if (key == rd.needle_id) {
                found = mid;
                break;
            }
            if (key < rd.needle_id) lo = mid + 1;
            else hi = mid;
        }
        if (found < 0) {
            rd.status = SWEC_ERR_NOT_FOUND;
            continue;
        }
        const int64_t offset = int64_t(be32(ex + found * 16 + 8)) * 8;  // Offset.ToActualOffset (offset_4bytes.go)
        int32_t size = int32_t(be32(ex + found * 16 + 12));
        if (journalled(rd.needle_id)) size = -1;  // TombstoneFileSize (FindNeedleFromEcx, ec_volume.go:419-429)
        rd.offset = offset;
        rd.size = size;
        if (size < 0) {  // Size.IsDeleted
            rd.status = SWEC_ERR_DELETED;
            continue;
        }
        // LocateEcShardNeedle passes GetActualSize(size) to LocateEcShardNeedleInterval, which applies
        // GetActualSize AGAIN (ec_volume.go:395,414): the reference reads a little past the record.  Kept,
        // because ReadEcShardNeedle returns len(bytes) and n.ReadBytes only parses the front.
        const int64_t want = needle_actual_size(needle_actual_size(size, version), version);
        if (!rd.buf || size_t(want) > rd.capacity) {
            rd.n_bytes = size_t(want);
            rd.status = SWEC_ERR_INVALID_ARG;
            continue;
        }
        std::vector<swec_interval> ivs(size_t(want / small) + 4);  // a record crosses at most size/1 MiB + 2 blocks
        const int niv = swec_locate_data(large, small, v->shard_dat_size, offset, want, k, ivs.data(), int(ivs.size()));
        if (niv < 0) {
            rd.status = niv;
            continue;
        }
        size_t pos = 0;
        for (int j = 0; j < niv && rd.status == SWEC_OK; j++) {
            int sid = 0;
            int64_t soff = 0;
            swec_interval_to_shard(&ivs[j], large, small, k, &sid, &soff);
            const size_t len = size_t(ivs[j].size);
            bool ok = false;
            if (v->shard_fd[size_t(sid)] >= 0) {  // readLocalEcShardInterval (store_ec.go:407-422): all or nothing
                const ssize_t got = pread(v->shard_fd[size_t(sid)], rd.buf + pos, len, off_t(soff));
                ok = got == ssize_t(len);
            }
            if (!ok) {  // recoverOneRemoteEcShardInterval, with "remote" = every other local shard file
                Recover rc;
                rc.read_idx = r;
                rc.buf_off = pos;
                rc.len = len;
                rc.shard = sid;
                rc.bufs.resize(size_t(total));
                rc.present.assign(size_t(total), 0);
                int have = 0;
                for (int i = 0; i < total; i++) {
                    if (i == sid || v->shard_fd[size_t(i)] < 0) continue;
                    rc.bufs[size_t(i)].resize(len);
                    if (pread(v->shard_fd[size_t(i)], rc.bufs[size_t(i)].data(), len, off_t(soff)) == ssize_t(len)) {
                        rc.present[size_t(i)] = 1;
                        have++;
                    } else {
                        rc.bufs[size_t(i)].clear();  // nRead != len(buf): not available
                    }
                }
                if (have < k) {
                    rd.status = SWEC_ERR_TOO_FEW_SHARDS;
                    set_last_error("cannot recover shard " + std::to_string(sid) + ": only " + std::to_string(have) +
                                   " shards available, need at least " + std::to_string(k));
                    break;
                }
                for (int i = 0; i < k; i++)  // ReconstructData fills every missing DATA shard
                    if (!rc.present[size_t(i)]) rc.bufs[size_t(i)].resize(len);
                recs.push_back(std::move(rc));
                rd.n_recovered_intervals++;
            }
            pos += len;
        }
        if (rd.status == SWEC_OK) rd.n_bytes = pos;
    }

    // ---- pass 2: every interval that needs the arithmetic, in ONE batched ReconstructData on the GPU
    if (!recs.empty()) {
        if (!v->enc) {
            const int rc = swec_encoder_new(v->k, v->m, v->device, &v->enc);
            if (rc) return rc;
        }
        std::vector<swec_reconstruct_item> items(recs.size());
        for (size_t j = 0; j < recs.size(); j++) {
            recs[j].ptrs.assign(size_t(total), nullptr);
            for (int i = 0; i < total; i++)
                if (!recs[j].bufs[size_t(i)].empty()) recs[j].ptrs[size_t(i)] = recs[j].bufs[size_t(i)].data();
            items[j].shards = recs[j].ptrs.data();
            items[j].present = recs[j].present.data();
            items[j].shard_len = recs[j].len;
            items[j].data_only = 1;
        }
        const int rc = swec_reconstruct_batch(v->enc, items.data(), int(items.size()));
        if (rc) return rc;
        for (const Recover& rcv : recs) {
            swec_needle_read& rd = reads[rcv.read_idx];
            if (rd.status != SWEC_OK) continue;
            memcpy(rd.buf + rcv.buf_off, rcv.bufs[size_t(rcv.shard)].data(), rcv.len);
        }
    }
    return SWEC_OK;
}

// EcVolume.ScrubLocal (ec_volume_scrub.go:27-118) minus the needle parse: ScrubIndex, then every live entry of .ecx is
// located (GetActualSize ONCE here, unlike the read path) and each of its chunks read from the local shard that holds it.
// A shard that is too short for a chunk, or cannot be read, is reported broken; chunks on shards that are not local are
// skipped like the reference skips remote ones.  The CRC of the record belongs to the storage engine's needle parser.
int swec_ec_volume_scrub_local(swec_ec_volume* v, int64_t* entries, uint32_t* broken_shards, int* n_broken, char* errors,
                               size_t errors_cap, int* n_errors) {
    if (!v || !entries || !n_broken || !n_errors || !broken_shards) return fail(SWEC_ERR_INVALID_ARG, "NULL argument");
    std::lock_guard<std::mutex> lock(v->mu);
    std::string text;
    int count = 0;
    auto add = [&](const std::string& m) {
        if (count++) text += "\n";
        text += m;
    };
    {  // ScrubIndex = idx.CheckIndexFile on the sealed index
        int64_t n = 0;
        int k2 = 0;
        std::vector<char> buf(size_t(1) << 20);
        const int rc = swec_check_index_file((v->index_base + ".ecx").c_str(), v->version, &n, buf.data(), buf.size(), &k2);
        if (rc) return rc;
        if (k2) add(buf.data()), count += k2 - 1;
    }
    const int k = v->k, total = v->k + v->m;
    const int64_t large = int64_t(1) << 30, small = int64_t(1) << 20;
    std::vector<int64_t> shard_size(size_t(total), -1);
    for (int i = 0; i < total; i++) {
        struct stat st;
        if (v->shard_fd[size_t(i)] >= 0 && fstat(v->shard_fd[size_t(i)], &st) == 0) shard_size[size_t(i)] = st.st_size;
    }
    std::vector<uint8_t> broken(size_t(total), 0), chunk;
    const uint8_t* ex = v->ecx_map;
    const int64_t n_entries = int64_t(v->ecx_bytes) / 16;
    int64_t walked = 0;
    for (int64_t e = 0; e < n_entries; e++) {
        walked++;
        const uint64_t id = be64(ex + e * 16);
        const int64_t offset = int64_t(be32(ex + e * 16 + 8)) * 8;
        const int32_t size = int32_t(be32(ex + e * 16 + 12));
        if (size == -1) continue;  // Size.IsTombstone
        const int64_t want = needle_actual_size(size, v->version);
        std::vector<swec_interval> ivs(size_t(std::max<int64_t>(want, 0) / small) + 4);
        const int niv = want > 0 ? swec_locate_data(large, small, v->shard_dat_size, offset, want, k, ivs.data(), int(ivs.size())) : 0;
        if (niv < 0) return niv;
        int64_t read = 0;
        for (int j = 0; j < niv; j++) {
            int sid = 0;
            int64_t soff = 0;
            swec_interval_to_shard(&ivs[size_t(j)], large, small, k, &sid, &soff);
            const int64_t ssize = ivs[size_t(j)].size;
            const std::string where = std::to_string(j + 1) + "/" + std::to_string(niv);
            if (v->shard_fd[size_t(sid)] < 0) {  // not local: skipped, counted as read
                read += ssize;
                continue;
            }
            if (soff + ssize > shard_size[size_t(sid)]) {
                broken[size_t(sid)] = 1;
                add("local shard " + std::to_string(sid) + " for needle " + std::to_string(id) + " is too short (" +
                    std::to_string(shard_size[size_t(sid)]) + "), cannot read chunk " + where);
                continue;
            }
            chunk.resize(size_t(ssize));
            const ssize_t got = pread(v->shard_fd[size_t(sid)], chunk.data(), size_t(ssize), off_t(soff));
            if (got < 0) {
                broken[size_t(sid)] = 1;
                add("failed to read chunk " + where + " for needle " + std::to_string(id) + " from local shard " + std::to_string(sid) +
                    " at offset " + std::to_string(soff) + ": " + strerror(errno));
                continue;
            }
            if (got != ssize) {
                broken[size_t(sid)] = 1;
                add("expected " + std::to_string(ssize) + " bytes for chunk " + where + " for needle " + std::to_string(id) +
                    " from local shard " + std::to_string(sid) + ", got " + std::to_string(got));
                continue;
            }
            read += got;
        }
        if (read != want) {  // the reference's walk stops here
            add("expected " + std::to_string(want) + " bytes for needle " + std::to_string(id) + ", got " + std::to_string(read));
            break;
        }
    }
    *entries = walked;
    *n_broken = 0;
    for (int i = 0; i < total; i++)
        if (broken[size_t(i)]) broken_shards[(*n_broken)++] = uint32_t(i);
    *n_errors = count;
    if (errors && errors_cap) {
        const size_t m = std::min(text.size(), errors_cap - 1);
        memcpy(errors, text.data(), m);
        errors[m] = 0;
    }
    return SWEC_OK;
}

// what NewEcVolume derived from .vif and the shard files (ec_volume.go:114-154,399-417)
int swec_ec_volume_info(swec_ec_volume* v, int* data_shards, int* parity_shards, int* needle_version, int64_t* shard_dat_size,
                        uint32_t* local_shard_bits) {
    if (!v) return fail(SWEC_ERR_INVALID_ARG, "NULL volume");
    std::lock_guard<std::mutex> lock(v->mu);
    if (data_shards) *data_shards = v->k;
    if (parity_shards) *parity_shards = v->m;
    if (needle_version) *needle_version = v->version;
    if (shard_dat_size) *shard_dat_size = v->shard_dat_size;
    if (local_shard_bits) {
        uint32_t bits = 0;
        for (size_t i = 0; i < v->shard_fd.size(); i++)
            if (v->shard_fd[i] >= 0) bits |= 1u << i;
        *local_shard_bits = bits;
    }
    return SWEC_OK;
}

// FileAndDeleteCount (ec_volume.go:330-349): entries of the sealed .ecx, and distinct journalled ids.
int swec_ec_volume_counts(swec_ec_volume* v, uint64_t* file_count, uint64_t* delete_count) {
    if (!v) return fail(SWEC_ERR_INVALID_ARG, "NULL volume");
    std::lock_guard<std::mutex> lock(v->mu);
    v->refresh_journal();
    if (file_count) *file_count = uint64_t(v->ecx_bytes / 16);
    if (delete_count) *delete_count = uint64_t(v->deleted.size());
    return SWEC_OK;
}

// DeleteNeedleFromEcx (ec_volume_delete.go:28-93): .ecx stays sealed; a runtime delete appends the id to the
// .ecj journal (the durable commit point: written, fsync'ed, truncated back on failure) and only then becomes
// visible to reads.  Unknown ids and ids that are already tombstoned / journalled are not errors.
int swec_ec_volume_delete_needle(swec_ec_volume* v, uint64_t needle_id) {
    if (!v) return fail(SWEC_ERR_INVALID_ARG, "NULL volume");
    std::lock_guard<std::mutex> lock(v->mu);
    const uint8_t* ex = v->ecx_map;
    int64_t lo = 0, hi = int64_t(v->ecx_bytes) / 16, found = -1;
    while (lo < hi) {
        const int64_t mid = (lo + hi) / 2;
        const uint64_t key = be64(ex + mid * 16);
        if (key == needle_id) {
            found = mid;
            break;
        }
        if (key < needle_id) lo = mid + 1;
        else hi = mid;
    }
    if (found < 0) return SWEC_OK;                                   // already gone
    if (int32_t(be32(ex + found * 16 + 12)) < 0) return SWEC_OK;     // folded into .ecx by an earlier rebuild
    // the in-memory set is authoritative between external changes of the file (refresh_journal notices those by
    // size / mtime / inode): an O(log n) membership test and an 8-byte append per delete, like the reference's map
    // check + append — not a re-read of the whole journal
    v->refresh_journal();
    if (std::binary_search(v->deleted.begin(), v->deleted.end(), needle_id)) return SWEC_OK;  // idempotent
    const std::string path = v->index_base + ".ecj";
    const int fd = open(path.c_str(), O_WRONLY | O_CREAT, 0644);
    if (fd < 0) return fail(SWEC_ERR_IO, "cannot open ec volume journal " + path + ": " + strerror(errno));
    struct stat st;
    if (fstat(fd, &st) != 0) {
        const int e = errno;
        close(fd);
        return fail(SWEC_ERR_IO, "stat ecj: " + std::string(strerror(e)));
    }
    uint8_t b[8];
    for (int i = 0; i < 8; i++) b[i] = uint8_t(needle_id >> (8 * (7 - i)));
    const bool ok = pwrite(fd, b, 8, st.st_size) == 8 && fsync(fd) == 0;
    const int e = errno;
    if (!ok) {
        if (ftruncate(fd, st.st_size) != 0) {}  // keep journal and in-memory state from drifting
        close(fd);
        return fail(SWEC_ERR_IO, "write ecj: " + std::string(strerror(e)));
    }
    struct stat after;
    const bool stat_ok = fstat(fd, &after) == 0;
    close(fd);
    v->ecj.append(reinterpret_cast<const char*>(b), 8);
    v->deleted.insert(std::upper_bound(v->deleted.begin(), v->deleted.end(), needle_id), needle_id);
    if (stat_ok) {
        v->ecj_size_seen = int64_t(after.st_size);
        v->ecj_mtime_ns_seen = int64_t(after.st_mtim.tv_sec) * 1000000000ll + after.st_mtim.tv_nsec;
        v->ecj_inode_seen = uint64_t(after.st_ino);
    } else {
        v->ecj_size_seen = -1;  // re-read on the next use
    }
    return SWEC_OK;
}

int swec_read_ec_needles(const char* data_base, const char* index_base, const char* const* additional_dirs,
                         int n_additional_dirs, swec_needle_read* reads, int n_reads, int device) {
    if (n_reads > 0 && !reads) return fail(SWEC_ERR_INVALID_ARG, "NULL argument");
    swec_ec_volume* v = nullptr;
    int rc = swec_ec_volume_open(data_base, index_base, additional_dirs, n_additional_dirs, device, &v);
    if (rc) return rc;
    rc = swec_ec_volume_read_needles(v, reads, n_reads);
    swec_ec_volume_close(v);
    return rc;
}

}  // extern "C"
