"""Host-side mirror of SeaweedFS's ``weed/storage/erasure_coding`` package surface for the RS hot
path, bound to libswec.so.  Names follow the Go package (Go spelling kept as aliases) so the parity
tests read like the reference's own tests:

    ECContext / NewDefaultECContext / CreateEncoder       ec_context.go:11-46
    Encoder.Encode / Reconstruct / ReconstructData        klauspost Encoder, call sites
                                                          ec_encoder.go:265,360; store_ec.go:551
    WriteEcFiles / generateEcFiles / RebuildEcFiles       ec_encoder.go:61-128,146-200
    WriteDatFile                                          ec_decoder.go:176-223
    LocateData / Interval.ToShardIdAndOffset              ec_locate.go:16-98

Buffers are numpy uint8 arrays (host) or raw device pointers (ints) for the ``*_device`` calls.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from ._native import STATUS, Interval, NeedleRead, ReconstructItem, SwecError, check, lib

DataShardsCount = 10                               # ec_encoder.go:20
ParityShardsCount = 4                              # ec_encoder.go:21
TotalShardsCount = DataShardsCount + ParityShardsCount
MaxShardCount = 32                                 # ec_encoder.go:23
ErasureCodingLargeBlockSize = 1024 * 1024 * 1024   # ec_encoder.go:25
ErasureCodingSmallBlockSize = 1024 * 1024          # ec_encoder.go:26


def _ptrs(arrs) -> C.Array:
    out = (C.c_void_p * len(arrs))()
    for i, a in enumerate(arrs):
        if a is None:
            out[i] = None
        elif isinstance(a, np.ndarray):
            out[i] = a.ctypes.data
        else:
            out[i] = int(a)
    return out


class Encoder:
    """reedsolomon.Encoder backed by the B200 engine (swec_encoder)."""

    def __init__(self, data_shards: int, parity_shards: int, device: int = 0):
        h = C.c_void_p()
        check(lib().swec_encoder_new(data_shards, parity_shards, device, C.byref(h)))
        self._h = h
        self.data_shards, self.parity_shards, self.device = data_shards, parity_shards, device

    def close(self) -> None:
        h, self._h = getattr(self, "_h", None), None
        if h and callable(lib):        # module globals are already cleared when this runs at interpreter exit
            lib().swec_encoder_free(h)

    __del__ = close

    @property
    def total_shards(self) -> int:
        return self.data_shards + self.parity_shards

    def matrix(self) -> np.ndarray:
        m = np.zeros((self.total_shards, self.data_shards), dtype=np.uint8)
        check(lib().swec_encoder_matrix(self._h, m.ctypes.data))
        return m

    def reconstruct_matrix(self, present, data_only: bool = False):
        pres = np.ascontiguousarray(np.asarray(present, dtype=np.uint8))
        inputs = (C.c_int * self.data_shards)()
        outputs = (C.c_int * self.total_shards)()
        n = C.c_int(0)
        rows = np.zeros((self.total_shards, self.data_shards), dtype=np.uint8)
        check(lib().swec_reconstruct_matrix(self._h, pres.ctypes.data, int(data_only), inputs, outputs,
                                            C.byref(n), rows.ctypes.data))
        return list(inputs), list(outputs[: n.value]), rows[: n.value].copy()

    # -- host buffers ---------------------------------------------------------------------------
    def _check_shards(self, shards, allow_missing: bool):
        if len(shards) != self.total_shards:
            raise SwecError(-1, f"need {self.total_shards} shards, got {len(shards)} (ErrTooFewShards)")
        n = None
        for s in shards:
            if s is None or (allow_missing and len(s) == 0):
                if not allow_missing:
                    raise SwecError(-1, "nil shard")
                continue
            if s.dtype != np.uint8 or not s.flags["C_CONTIGUOUS"]:
                raise SwecError(-1, "shards must be contiguous uint8 arrays")
            if n is None:
                n = s.shape[0]
            elif s.shape[0] != n:
                raise SwecError(-6, "shards are of different sizes (ErrShardSize)")
        if not n:
            raise SwecError(-1, "no shard data (ErrShardNoData)")
        return n

    def encode(self, shards: list[np.ndarray]) -> None:
        """Encode(shards [][]byte): parity slices (last m) are overwritten in place."""
        n = self._check_shards(shards, allow_missing=False)
        check(lib().swec_encode(self._h, _ptrs(shards), n))

    def reconstruct(self, shards: list, data_only: bool = False) -> None:
        """Reconstruct(shards): None / empty entries are missing and are replaced by new arrays."""
        n = self._check_shards(shards, allow_missing=True)
        present = np.array([s is not None and len(s) > 0 for s in shards], dtype=np.uint8)
        if present.sum() < self.data_shards:
            raise SwecError(-2, "too few shards given (ErrTooFewShards)")
        for i in range(self.total_shards):
            if not present[i] and (i < self.data_shards or not data_only):
                shards[i] = np.zeros(n, dtype=np.uint8)
        bufs = [s if s is not None and len(s) else None for s in shards]
        check(lib().swec_reconstruct(self._h, _ptrs(bufs), present.ctypes.data, n, int(data_only)))

    def reconstruct_data(self, shards: list) -> None:
        self.reconstruct(shards, data_only=True)

    def reconstruct_batch(self, batch: list[list], data_only: bool = True) -> None:
        """Many Reconstruct/ReconstructData calls in one crossing (batched degraded reads): each
        element is a shards list as for reconstruct(); missing entries are filled in place."""
        items = (ReconstructItem * len(batch))()
        keep = []
        for j, shards in enumerate(batch):
            n = self._check_shards(shards, allow_missing=True)
            present = np.array([s is not None and len(s) > 0 for s in shards], dtype=np.uint8)
            for i in range(self.total_shards):
                if not present[i] and (i < self.data_shards or not data_only):
                    shards[i] = np.zeros(n, dtype=np.uint8)
            ptrs = _ptrs([s if s is not None and len(s) else None for s in shards])
            keep.append((ptrs, present))
            items[j].shards = C.cast(ptrs, C.POINTER(C.c_void_p))
            items[j].present = present.ctypes.data_as(C.POINTER(C.c_uint8))
            items[j].shard_len = n
            items[j].data_only = int(data_only)
        check(lib().swec_reconstruct_batch(self._h, items, len(batch)))

    def verify(self, shards: list[np.ndarray]) -> bool:
        n = self._check_shards(shards, allow_missing=False)
        ok = C.c_int(0)
        check(lib().swec_verify(self._h, _ptrs(shards), n, C.byref(ok)))
        return bool(ok.value)

    # -- device buffers (raw pointers), asynchronous on `stream` ----------------------------------
    def encode_device(self, data_ptrs, parity_ptrs, shard_len: int, stream: int = 0) -> None:
        check(lib().swec_encode_device(self._h, _ptrs(data_ptrs), _ptrs(parity_ptrs), shard_len, stream))

    def reconstruct_device(self, shard_ptrs, present, shard_len: int, data_only: bool = False, stream: int = 0) -> None:
        pres = np.ascontiguousarray(np.asarray(present, dtype=np.uint8))
        check(lib().swec_reconstruct_device(self._h, _ptrs(shard_ptrs), pres.ctypes.data, shard_len,
                                            int(data_only), stream))

    def apply_device(self, rows, in_ptrs, out_ptrs, shard_len: int, stream: int = 0) -> None:
        """out[p] = XOR_i rows[p][i] ⊗ in[i] for an arbitrary matrix (device pointers)."""
        rows = np.ascontiguousarray(rows, dtype=np.uint8)
        check(lib().swec_apply_device(self._h, rows.shape[0], rows.shape[1], rows.ctypes.data, _ptrs(in_ptrs),
                                      _ptrs(out_ptrs), shard_len, stream))

    def encode_volume_device(self, dat_ptr: int, dat_size: int, parity_ptrs, stream: int = 0,
                             large_block: int = ErasureCodingLargeBlockSize,
                             small_block: int = ErasureCodingSmallBlockSize) -> None:
        check(lib().swec_encode_volume_device(self._h, dat_ptr, dat_size, large_block, small_block,
                                              _ptrs(parity_ptrs), stream))

    def extract_data_shard_device(self, dat_ptr: int, dat_size: int, shard_id: int, out_ptr: int, stream: int = 0,
                                  large_block: int = ErasureCodingLargeBlockSize,
                                  small_block: int = ErasureCodingSmallBlockSize) -> None:
        check(lib().swec_extract_data_shard_device(self._h, dat_ptr, dat_size, large_block, small_block,
                                                   shard_id, out_ptr, stream))

    def write_dat_device(self, data_shard_ptrs, dat_size: int, dat_out_ptr: int, stream: int = 0,
                         large_block: int = ErasureCodingLargeBlockSize,
                         small_block: int = ErasureCodingSmallBlockSize) -> None:
        """WriteDatFile on device memory (ec_decoder.go:176-223): un-stripe the k data shards into the volume image."""
        check(lib().swec_write_dat_device(self._h, _ptrs(data_shard_ptrs), dat_size, large_block, small_block,
                                          dat_out_ptr, stream))

    def synchronize(self, stream: int = 0) -> None:
        check(lib().swec_stream_synchronize(self._h, stream))

    # Go spellings
    Encode, Reconstruct, ReconstructData, Verify = encode, reconstruct, reconstruct_data, verify


class EncoderGroup:
    """Several Encoder handles (one per GPU) behind the Encoder interface: every host-buffer call is split
    by byte-column range across the handles (swec_encode_multi / swec_reconstruct_multi)."""

    def __init__(self, data_shards: int, parity_shards: int, devices: list[int]):
        self.encoders = [Encoder(data_shards, parity_shards, d) for d in devices]
        self.data_shards, self.parity_shards = data_shards, parity_shards
        self._arr = (C.c_void_p * len(self.encoders))(*[e._h for e in self.encoders])

    @property
    def total_shards(self) -> int:
        return self.data_shards + self.parity_shards

    def close(self) -> None:
        for e in self.encoders:
            e.close()
        self.encoders = []

    def encode(self, shards: list[np.ndarray]) -> None:
        n = self.encoders[0]._check_shards(shards, allow_missing=False)
        check(lib().swec_encode_multi(self._arr, len(self.encoders), _ptrs(shards), n))

    def reconstruct(self, shards: list, data_only: bool = False) -> None:
        e0 = self.encoders[0]
        n = e0._check_shards(shards, allow_missing=True)
        present = np.array([s is not None and len(s) > 0 for s in shards], dtype=np.uint8)
        if present.sum() < self.data_shards:
            raise SwecError(-2, "too few shards given (ErrTooFewShards)")
        for i in range(self.total_shards):
            if not present[i] and (i < self.data_shards or not data_only):
                shards[i] = np.zeros(n, dtype=np.uint8)
        bufs = [s if s is not None and len(s) else None for s in shards]
        check(lib().swec_reconstruct_multi(self._arr, len(self.encoders), _ptrs(bufs), present.ctypes.data, n,
                                           int(data_only)))

    Encode, Reconstruct = encode, reconstruct


@dataclass
class ECContext:
    """ec_context.go:11-46"""
    DataShards: int = DataShardsCount
    ParityShards: int = ParityShardsCount
    Collection: str = ""
    VolumeId: int = 0
    device: int = 0

    def Total(self) -> int:
        return self.DataShards + self.ParityShards

    def CreateEncoder(self) -> Encoder:
        return Encoder(self.DataShards, self.ParityShards, self.device)

    def ToExt(self, shard_index: int) -> str:
        return ".ec%02d" % shard_index

    def String(self) -> str:
        return "%d+%d (total: %d)" % (self.DataShards, self.ParityShards, self.Total())


def NewDefaultECContext(collection: str = "", volume_id: int = 0, device: int = 0) -> ECContext:
    return ECContext(DataShardsCount, ParityShardsCount, collection, volume_id, device)


def ToExt(ec_index: int) -> str:
    return ".ec%02d" % ec_index


# ---- file level ---------------------------------------------------------------------------------

def generate_ec_files(base_file_name: str, buffer_size: int, large_block_size: int, small_block_size: int,
                      ctx: ECContext | None = None) -> None:
    ctx = ctx or NewDefaultECContext()
    check(lib().swec_generate_ec_files(base_file_name.encode(), buffer_size, large_block_size, small_block_size,
                                       ctx.DataShards, ctx.ParityShards, ctx.device))


def write_ec_files(base_file_name: str, ctx: ECContext | None = None) -> None:
    generate_ec_files(base_file_name, 256 * 1024, ErasureCodingLargeBlockSize, ErasureCodingSmallBlockSize, ctx)


def rebuild_ec_files(base_file_name: str, additional_dirs: list[str] | None = None,
                     ctx: ECContext | None = None, device: int = 0) -> list[int]:
    """RebuildEcFiles (ctx None ⇒ ratio from .vif or 10+4) / RebuildEcFilesWithContext."""
    dirs = [d.encode() for d in (additional_dirs or [])]
    arr = (C.c_char_p * max(1, len(dirs)))(*dirs) if dirs else None
    ids = (C.c_uint32 * MaxShardCount)()
    n = C.c_int(0)
    k, m, dev = (ctx.DataShards, ctx.ParityShards, ctx.device) if ctx else (0, 0, device)
    check(lib().swec_rebuild_ec_files(base_file_name.encode(), arr, len(dirs), k, m, dev, ids, C.byref(n)))
    return list(ids[: n.value])


def verify_ec_files(base_file_name: str, additional_dirs: list[str] | None = None, ctx: ECContext | None = None,
                    device: int = 0):
    """Parity scrub of a shard set: returns (ok, mismatching 16-byte vectors per parity shard)."""
    dirs = [d.encode() for d in (additional_dirs or [])]
    arr = (C.c_char_p * max(1, len(dirs)))(*dirs) if dirs else None
    k, m, dev = (ctx.DataShards, ctx.ParityShards, ctx.device) if ctx else (0, 0, device)
    bad = (C.c_uint64 * MaxShardCount)()
    ok = C.c_int(0)
    check(lib().swec_verify_ec_files(base_file_name.encode(), arr, len(dirs), k, m, dev, bad, C.byref(ok)))
    nm = ctx.ParityShards if ctx else ParityShardsCount
    return bool(ok.value), list(bad[:nm])


def write_dat_file(base_file_name: str, dat_file_size: int, shard_file_names: list[str],
                   data_shards: int = DataShardsCount, large_block: int = ErasureCodingLargeBlockSize,
                   small_block: int = ErasureCodingSmallBlockSize) -> None:
    names = (C.c_char_p * len(shard_file_names))(*[s.encode() for s in shard_file_names])
    check(lib().swec_write_dat_file(base_file_name.encode(), dat_file_size, names, data_shards, large_block, small_block))


# ---- whole-volume operations (the file work of the three EC gRPC handlers) --------------------------

def _dirs(additional_dirs):
    dirs = [d.encode() for d in (additional_dirs or [])]
    return ((C.c_char_p * max(1, len(dirs)))(*dirs) if dirs else None), len(dirs)


def volume_ec_shards_generate(data_base_file_name: str, index_base_file_name: str | None = None,
                              needle_version: int = 0, expire_at_sec: int = 0, device: int = 0) -> None:
    """VolumeEcShardsGenerate (volume_grpc_erasure_coding.go:43-146): .ecx first, shards, .vif; cleanup on error."""
    check(lib().swec_ec_shards_generate(data_base_file_name.encode(),
                                        (index_base_file_name or "").encode(), needle_version, expire_at_sec, device))


def volume_ec_shards_rebuild(data_base_file_name: str, index_base_file_name: str | None = None,
                             additional_dirs: list[str] | None = None, device: int = 0) -> list[int]:
    """VolumeEcShardsRebuild (volume_grpc_erasure_coding.go:149-225): RebuildEcFiles + RebuildEcxFile."""
    arr, n = _dirs(additional_dirs)
    ids = (C.c_uint32 * MaxShardCount)()
    cnt = C.c_int(0)
    check(lib().swec_ec_shards_rebuild(data_base_file_name.encode(), (index_base_file_name or "").encode(),
                                       arr, n, device, ids, C.byref(cnt)))
    return list(ids[: cnt.value])


def volume_ec_shards_to_volume(data_base_file_name: str, index_base_file_name: str | None = None,
                               additional_dirs: list[str] | None = None) -> int:
    """VolumeEcShardsToVolume (volume_grpc_erasure_coding.go:578-668): shards + .ecx/.ecj → .dat + .idx.
    Returns the .dat size; raises SwecError(SWEC_ERR_NO_LIVE_NEEDLES) for an all-deleted volume."""
    arr, n = _dirs(additional_dirs)
    size = C.c_int64(0)
    check(lib().swec_ec_shards_to_volume(data_base_file_name.encode(), (index_base_file_name or "").encode(),
                                         arr, n, C.byref(size)))
    return int(size.value)


def _needle_reads(needle_ids, capacity, sizes=None):
    """capacity: fixed bytes per needle, or None with `sizes` (exact bytes per needle from a sizing pass)."""
    reads = (NeedleRead * len(needle_ids))()
    caps = [capacity] * len(needle_ids) if sizes is None else list(sizes)
    starts = np.concatenate([[0], np.cumsum(caps)]).astype(np.int64)
    arena = np.empty(max(1, int(starts[-1])), dtype=np.uint8)
    for j, (r, nid) in enumerate(zip(reads, needle_ids)):
        r.needle_id, r.buf, r.capacity = nid, arena.ctypes.data + int(starts[j]), caps[j]
    return reads, arena, starts


def _needle_results(reads, arena, starts):
    return [{"id": r.needle_id, "status": STATUS.get(r.status, str(r.status)), "offset": r.offset, "size": r.size,
             "bytes": arena[int(starts[j]): int(starts[j]) + r.n_bytes] if r.status == 0 else None,
             "n_bytes": r.n_bytes, "recovered_intervals": r.n_recovered_intervals} for j, r in enumerate(reads)]


def _read_needles(call, needle_ids, capacity):
    """capacity None ⇒ two passes: a sizing pass with zero-capacity buffers (every live needle answers
    SWEC_ERR_INVALID_ARG with the bytes it needs, nothing is read), then the real one into an exact arena."""
    if capacity is None:
        reads, arena, starts = _needle_reads(needle_ids, 0)
        check(call(reads, len(needle_ids)))
        sizes = [r.n_bytes if r.status == -1 else 0 for r in reads]
        reads, arena, starts = _needle_reads(needle_ids, None, sizes)
    else:
        reads, arena, starts = _needle_reads(needle_ids, capacity)
    check(call(reads, len(needle_ids)))
    return _needle_results(reads, arena, starts)


class EcVolume:
    """A mounted EC volume (erasure_coding.EcVolume, ec_volume.go:36-160) for the read path: shard files open,
    .ecx loaded, the encoder behind degraded reads kept warm.  ReadEcShardNeedles = Store.ReadEcShardNeedle
    (store_ec.go:252-355) for many needles in one call."""

    def __init__(self, data_base_file_name: str, index_base_file_name: str | None = None,
                 additional_dirs: list[str] | None = None, device: int = 0):
        arr, n = _dirs(additional_dirs)
        h = C.c_void_p()
        check(lib().swec_ec_volume_open(data_base_file_name.encode(), (index_base_file_name or "").encode(), arr, n,
                                        device, C.byref(h)))
        self._h = h

    def read_needles(self, needle_ids: list[int], capacity: int | None = None):
        return _read_needles(lambda reads, n: lib().swec_ec_volume_read_needles(self._h, reads, n), needle_ids, capacity)

    def delete_needle(self, needle_id: int) -> None:
        """DeleteNeedleFromEcx (ec_volume_delete.go:28-93)"""
        check(lib().swec_ec_volume_delete_needle(self._h, needle_id))

    DeleteNeedleFromEcx = delete_needle

    def info(self) -> dict:
        k, m, ver, bits, sds = C.c_int(0), C.c_int(0), C.c_int(0), C.c_uint32(0), C.c_int64(0)
        check(lib().swec_ec_volume_info(self._h, C.byref(k), C.byref(m), C.byref(ver), C.byref(sds), C.byref(bits)))
        return {"data_shards": k.value, "parity_shards": m.value, "version": ver.value, "shard_dat_size": sds.value,
                "local_shards": [i for i in range(32) if bits.value >> i & 1]}

    def file_and_delete_count(self) -> tuple[int, int]:
        """FileAndDeleteCount (ec_volume.go:330-349)"""
        f, d = C.c_uint64(0), C.c_uint64(0)
        check(lib().swec_ec_volume_counts(self._h, C.byref(f), C.byref(d)))
        return int(f.value), int(d.value)

    FileAndDeleteCount = file_and_delete_count

    def scrub_local(self):
        """ScrubLocal (ec_volume_scrub.go:27-118): (entries walked, broken shard ids, findings)."""
        n, nb, ne = C.c_int64(0), C.c_int(0), C.c_int(0)
        broken = (C.c_uint32 * MaxShardCount)()
        buf = C.create_string_buffer(1 << 20)
        check(lib().swec_ec_volume_scrub_local(self._h, C.byref(n), broken, C.byref(nb), buf, len(buf), C.byref(ne)))
        return int(n.value), list(broken[: nb.value]), (buf.value.decode().split("\n") if ne.value else [])

    ScrubLocal = scrub_local

    def close(self) -> None:
        h, self._h = getattr(self, "_h", None), None
        if h and callable(lib):
            lib().swec_ec_volume_close(h)

    __del__ = close
    ReadEcShardNeedles = read_needles


def read_ec_shard_needles(data_base_file_name: str, needle_ids: list[int], index_base_file_name: str | None = None,
                          additional_dirs: list[str] | None = None, device: int = 0, capacity: int | None = None):
    """One-shot form: mount, read, unmount.  Returns one dict per id with status name, offset, size, the raw
    record bytes and how many intervals had to be reconstructed."""
    arr, n = _dirs(additional_dirs)
    return _read_needles(lambda reads, cnt: lib().swec_read_ec_needles(
        data_base_file_name.encode(), (index_base_file_name or "").encode(), arr, n, reads, cnt, device),
        needle_ids, capacity)


ReadEcShardNeedles = read_ec_shard_needles
VolumeEcShardsGenerate, VolumeEcShardsRebuild, VolumeEcShardsToVolume = (
    volume_ec_shards_generate, volume_ec_shards_rebuild, volume_ec_shards_to_volume)


# ---- index files (.idx / .ecx / .ecj) --------------------------------------------------------------

def write_sorted_file_from_idx(base_file_name: str, ext: str = ".ecx") -> None:
    """WriteSortedFileFromIdx (ec_encoder.go:31-58)"""
    check(lib().swec_write_sorted_file_from_idx(base_file_name.encode(), ext.encode()))


def rebuild_ecx_file(base_file_name: str) -> None:
    """RebuildEcxFile (ec_volume_delete.go:95-142)"""
    check(lib().swec_rebuild_ecx_file(base_file_name.encode()))


def write_idx_file_from_ec_index(base_file_name: str) -> None:
    """WriteIdxFileFromEcIndex (ec_decoder.go:35-60)"""
    check(lib().swec_write_idx_file_from_ec_index(base_file_name.encode()))


def check_index_file(path: str, needle_version: int = 3) -> tuple[int, list[str]]:
    """idx.CheckIndexFile (weed/storage/idx/check.go:36-111): (entries processed, findings)."""
    n, k = C.c_int64(0), C.c_int(0)
    buf = C.create_string_buffer(1 << 20)
    check(lib().swec_check_index_file(path.encode(), needle_version, C.byref(n), buf, len(buf), C.byref(k)))
    text = buf.value.decode()
    return int(n.value), (text.split("\n") if k.value else [])


CheckIndexFile = check_index_file


def has_live_needles(index_base_file_name: str) -> bool:
    v = C.c_int(0)
    check(lib().swec_has_live_needles(index_base_file_name.encode(), C.byref(v)))
    return bool(v.value)


def find_dat_file_size(data_base_file_name: str, index_base_file_name: str) -> int:
    v = C.c_int64(0)
    check(lib().swec_find_dat_file_size(data_base_file_name.encode(), index_base_file_name.encode(), C.byref(v)))
    return int(v.value)


WriteSortedFileFromIdx, RebuildEcxFile, WriteIdxFileFromEcIndex = (write_sorted_file_from_idx, rebuild_ecx_file,
                                                                    write_idx_file_from_ec_index)
HasLiveNeedles, FindDatFileSize = has_live_needles, find_dat_file_size
WriteEcFiles, WriteEcFilesWithContext = write_ec_files, write_ec_files
generateEcFiles, RebuildEcFiles, WriteDatFile = generate_ec_files, rebuild_ec_files, write_dat_file


# ---- layout --------------------------------------------------------------------------------------

def expected_shard_size(dat_size: int, data_shards: int = DataShardsCount,
                        large_block: int = ErasureCodingLargeBlockSize,
                        small_block: int = ErasureCodingSmallBlockSize) -> int:
    return int(lib().swec_expected_shard_size(dat_size, data_shards, large_block, small_block))


def locate_data(large_block_length: int, small_block_length: int, shard_dat_size: int, offset: int, size: int,
                data_shards: int = DataShardsCount):
    """LocateData → list of (BlockIndex, InnerBlockOffset, Size, IsLargeBlock, LargeBlockRowsCount)."""
    cap = 4 + int(size // max(1, small_block_length)) + 2
    buf = (Interval * cap)()
    n = lib().swec_locate_data(large_block_length, small_block_length, shard_dat_size, offset, size,
                               data_shards, buf, cap)
    if n < 0:
        check(n)
    return [(b.block_index, b.inner_block_offset, b.size, bool(b.is_large_block), b.large_block_rows_count)
            for b in buf[:n]]


def interval_to_shard(iv, large_block: int, small_block: int, data_shards: int = DataShardsCount):
    """Interval.ToShardIdAndOffset"""
    c = Interval(iv[0], int(iv[3]), iv[1], iv[2], iv[4], 0)
    sid, off = C.c_int(0), C.c_int64(0)
    lib().swec_interval_to_shard(C.byref(c), large_block, small_block, data_shards, C.byref(sid), C.byref(off))
    return sid.value, off.value


LocateData = locate_data
