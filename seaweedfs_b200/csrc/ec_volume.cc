// seaweedfs_b200/csrc/ec_volume.cc — the volume-server-local bodies of the three gRPC handlers that
// drive the RS path, as single C-ABI calls:
//   swec_ec_shards_generate    VolumeEcShardsGenerate   weed/server/volume_grpc_erasure_coding.go:43-146
//   swec_ec_shards_rebuild     VolumeEcShardsRebuild    weed/server/volume_grpc_erasure_coding.go:149-225
//   swec_ec_shards_to_volume   VolumeEcShardsToVolume   weed/server/volume_grpc_erasure_coding.go:578-668
// Everything the handlers do to FILES is here, in the reference's order (.ecx before the shards, the
// .dat size snapshot before encoding, .vif last, partial outputs removed on any error); what they do
// to the server's in-memory state (volume lookup, maintenance mode, disk-location scan, compaction)
// stays in Go.  The shard arithmetic runs on the GPU through swec_generate_ec_files /
// swec_rebuild_ec_files; the index and .vif work is host-only.
#include <errno.h>
#include <fcntl.h>
#include <libgen.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cinttypes>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "engine.h"

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

}  // extern "C"
