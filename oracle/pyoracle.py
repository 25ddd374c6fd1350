"""oracle/pyoracle.py — ctypes binding of liboracle.so (CPU ORACLE, test infrastructure only).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package seaweedfs_b200 never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
_lib = None

u8p = C.POINTER(C.c_uint8)


def build(quiet: bool = True) -> None:
    """Compile liboracle.so and (when /root/reference exists) oracle/_ref/."""
    subprocess.run(["make", "-C", HERE] + (["-s"] if quiet else []), check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.orc_log_table.restype = u8p
        L.orc_exp_table.restype = u8p
        L.orc_mul_table.restype = u8p
        L.orc_mul.restype = C.c_uint8
        L.orc_mul.argtypes = [C.c_uint8, C.c_uint8]
        L.orc_div.restype = C.c_uint8
        L.orc_div.argtypes = [C.c_uint8, C.c_uint8]
        L.orc_exp.restype = C.c_uint8
        L.orc_exp.argtypes = [C.c_uint8, C.c_size_t]
        L.orc_mul_slice.argtypes = [C.c_uint8, C.c_void_p, C.c_void_p, C.c_size_t]
        L.orc_mul_slice_xor.argtypes = [C.c_uint8, C.c_void_p, C.c_void_p, C.c_size_t]
        L.orc_mul_table_half.argtypes = [C.c_uint8, C.c_void_p, C.c_void_p]
        L.orc_matrix_invert.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_build_matrix.argtypes = [C.c_int, C.c_int, C.c_void_p]
        L.orc_encode.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_size_t]
        L.orc_reconstruct.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
        L.orc_verify.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_size_t]
        L.orc_decode_matrix.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_expected_shard_size.restype = C.c_int64
        L.orc_expected_shard_size.argtypes = [C.c_int64, C.c_int, C.c_int64, C.c_int64]
        L.orc_encode_dat_image.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int64,
                                           C.c_int64, C.c_int64, C.c_void_p]
        L.orc_write_dat_image.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int64, C.c_int64, C.c_void_p]
        L.orc_generate_ec_files.argtypes = [C.c_char_p, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_int]
        L.orc_rebuild_ec_files.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_synth_fill.argtypes = [C.c_void_p, C.c_uint64, C.c_size_t, C.c_uint64]
        L.orc_ref_load.argtypes = [C.c_char_p]
        L.orc_ref_isa.restype = C.c_char_p
        L.orc_cpu_apply_mt.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_size_t, C.c_int, C.c_size_t]

        L.orc_cpu_bench.restype = C.c_double
        L.orc_cpu_bench.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_size_t]

        L.orc_volume_digests.argtypes = [C.c_int, C.c_int64, C.c_uint64, C.c_int, C.c_int, C.c_int64, C.c_int64,
                                         C.c_int, C.c_void_p]

        L.orc_generate_ec_files_mt.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int, C.c_int]
        L.orc_generate_ec_files_simd.argtypes = [C.c_char_p, C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_int]

        class Interval(C.Structure):
            _fields_ = [("block_index", C.c_int), ("inner_block_offset", C.c_int64),
                        ("size", C.c_int64), ("is_large_block", C.c_int),
                        ("large_block_rows_count", C.c_int)]

        L.Interval = Interval
        L.orc_locate_data.argtypes = [C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int,
                                      C.POINTER(Interval), C.c_int]
        _lib = L
    return _lib


def _ptr_array(arrs):
    return (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])


def table(name: str) -> np.ndarray:
    L = lib()
    n = {"log": 256, "exp": 510, "mul": 65536}[name]
    p = getattr(L, f"orc_{name}_table")()
    return np.ctypeslib.as_array(p, shape=(n,)).copy()


def build_matrix(k: int, total: int) -> np.ndarray:
    out = np.zeros((total, k), dtype=np.uint8)
    rc = lib().orc_build_matrix(k, total, out.ctypes.data)
    if rc:
        raise ValueError(f"orc_build_matrix rc={rc}")
    return out


def matrix_invert(m: np.ndarray) -> np.ndarray:
    m = np.ascontiguousarray(m, dtype=np.uint8)
    out = np.zeros_like(m)
    rc = lib().orc_matrix_invert(m.ctypes.data, m.shape[0], out.ctypes.data)
    if rc:
        raise ValueError("singular matrix")
    return out


def encode(k: int, m: int, data: list[np.ndarray]) -> list[np.ndarray]:
    n = data[0].shape[0]
    data = [np.ascontiguousarray(d, dtype=np.uint8) for d in data]
    par = [np.zeros(n, dtype=np.uint8) for _ in range(m)]
    rc = lib().orc_encode(k, m, _ptr_array(data + par), n)
    if rc:
        raise ValueError(f"orc_encode rc={rc}")
    return par


def reconstruct(k: int, m: int, shards: list, data_only: bool = False) -> list[np.ndarray]:
    n = next(s for s in shards if s is not None).shape[0]
    present = np.array([s is not None for s in shards], dtype=np.uint8)
    bufs = [np.ascontiguousarray(s, dtype=np.uint8).copy() if s is not None else np.zeros(n, dtype=np.uint8)
            for s in shards]
    rc = lib().orc_reconstruct(k, m, _ptr_array(bufs), present.ctypes.data, n, int(data_only))
    if rc:
        raise ValueError(f"orc_reconstruct rc={rc}")
    return bufs


def expected_shard_size(dat_size: int, k: int = 10, large: int = 1 << 30, small: int = 1 << 20) -> int:
    return int(lib().orc_expected_shard_size(dat_size, k, large, small))


def encode_dat_image(dat: np.ndarray, k=10, m=4, buffer_size=256 * 1024, large=1 << 30, small=1 << 20):
    dat = np.ascontiguousarray(dat, dtype=np.uint8)
    sz = expected_shard_size(dat.shape[0], k, large, small)
    shards = [np.zeros(sz, dtype=np.uint8) for _ in range(k + m)]
    rc = lib().orc_encode_dat_image(dat.ctypes.data, dat.shape[0], k, m, buffer_size, large, small,
                                    _ptr_array(shards))
    if rc:
        raise ValueError(f"orc_encode_dat_image rc={rc}")
    return shards


def write_dat_image(shards: list[np.ndarray], dat_size: int, k=10, large=1 << 30, small=1 << 20) -> np.ndarray:
    dat = np.zeros(dat_size, dtype=np.uint8)
    lib().orc_write_dat_image(dat.ctypes.data, dat_size, k, large, small, _ptr_array(shards[:k]))
    return dat


def locate_data(large, small, shard_dat_size, offset, size, k=10):
    L = lib()
    buf = (L.Interval * 64)()
    n = L.orc_locate_data(large, small, shard_dat_size, offset, size, k, buf, 64)
    if n < 0:
        raise ValueError("too many intervals")
    return [(b.block_index, b.inner_block_offset, b.size, bool(b.is_large_block), b.large_block_rows_count)
            for b in buf[:n]]


def generate_ec_files(base: str, buffer_size=256 * 1024, large=1 << 30, small=1 << 20, k=10, m=4) -> int:
    return lib().orc_generate_ec_files(base.encode(), buffer_size, large, small, k, m)


def rebuild_ec_files(base: str, k=10, m=4):
    ids = (C.c_uint32 * 32)()
    n = C.c_int(0)
    rc = lib().orc_rebuild_ec_files(base.encode(), k, m, ids, C.byref(n))
    return rc, list(ids[: n.value])


def synth(byte_offset: int, n: int, seed: int) -> np.ndarray:
    out = np.zeros(n, dtype=np.uint8)
    lib().orc_synth_fill(out.ctypes.data, byte_offset, n, seed)
    return out


def ref_available() -> bool:
    return os.path.exists(os.path.join(REF_DIR, "libref_rs_ssse3.so")) and \
        lib().orc_ref_load(REF_DIR.encode()) == 0


def ref_isa() -> str:
    return lib().orc_ref_isa().decode()


def gfni_level() -> int:
    return int(lib().orc_cpu_has_gfni())


def cpu_apply(kind: int, rows: np.ndarray, inputs: list[np.ndarray], outputs: list[np.ndarray],
              threads: int = 1, batch: int = 256 * 1024) -> None:
    """kind 0 = reference's compiled C kernel (oracle/_ref), kind 1 = AVX-512/AVX2+GFNI port."""
    rows = np.ascontiguousarray(rows, dtype=np.uint8)
    r, k = rows.shape
    if kind == 0 and not ref_available():
        raise RuntimeError("oracle/_ref not built")
    rc = lib().orc_cpu_apply_mt(kind, k, r, rows.ctypes.data, _ptr_array(inputs), _ptr_array(outputs),
                                inputs[0].shape[0], threads, batch)
    if rc:
        raise RuntimeError(f"cpu baseline kind {kind} unavailable on this CPU")


def cpu_bench(kind: int, rows: np.ndarray, bytes_per_shard: int, threads: int, passes: int,
              batch: int = 256 * 1024) -> float:
    """Whole-box CPU throughput (input GB/s): NUMA-local per-thread buffers, pinned threads, barrier
    start; every thread runs `passes` passes over its own k×bytes_per_shard inputs."""
    rows = np.ascontiguousarray(rows, dtype=np.uint8)
    r, k = rows.shape
    if kind == 0 and not ref_available():
        return -1.0
    return float(lib().orc_cpu_bench(kind, k, r, rows.ctypes.data, bytes_per_shard, threads, passes, batch))


def generate_ec_files_simd(base: str, kind: int, buffer_size=256 * 1024, large=1 << 30, small=1 << 20, k=10, m=4) -> int:
    """The reference-shaped serial file walk with SIMD Encode (kind 0 = reference C kernel, 1 = GFNI port)."""
    if kind == 0 and not ref_available():
        return -38
    return lib().orc_generate_ec_files_simd(base.encode(), kind, buffer_size, large, small, k, m)


def np_digest(arr: np.ndarray) -> int:
    """The digest function of the device's swec_digest_kernel, restated in numpy:
    Σ_j splitmix64_at(word_j, j) mod 2^64 over little-endian 8-byte words, tail zero-padded."""
    pad = (-len(arr)) % 8
    w = np.concatenate([arr, np.zeros(pad, dtype=np.uint8)]).view("<u8").astype(np.uint64)
    with np.errstate(over="ignore"):
        j = np.arange(len(w), dtype=np.uint64)
        z = w + (j + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
        return int(z.sum(dtype=np.uint64))


def best_cpu_kind() -> int:
    """0 = the reference's compiled C kernel (oracle/_ref), else 1 = GFNI port, else 2 = scalar tables."""
    if ref_available():
        return 0
    return 1 if gfni_level() else 2


def volume_digests(dat_size: int, seed: int, k=10, m=4, large=1 << 30, small=1 << 20, threads: int | None = None,
                   kind: int | None = None) -> list[int]:
    """Expected digests of all k+m shards of the seeded synthetic volume (orc_volume_digests): the whole
    volume, regenerated and encoded on the CPU chunk by chunk, nothing held in memory."""
    out = (C.c_uint64 * (k + m))()
    kind = best_cpu_kind() if kind is None else kind
    if kind == 0 and not ref_available():       # loads oracle/_ref on first use
        raise RuntimeError("oracle/_ref not built")
    rc = lib().orc_volume_digests(kind, dat_size, seed, k, m, large, small, threads or os.cpu_count() or 1, out)
    if rc:
        raise RuntimeError(f"orc_volume_digests rc={rc}")
    return [int(v) for v in out]


def generate_ec_files_mt(base: str, kind: int | None = None, threads: int | None = None, large=1 << 30, small=1 << 20,
                         k=10, m=4) -> int:
    """generateEcFiles on the CPU at the GPU pipeline's schedule (orc_generate_ec_files_mt): the same-schedule CPU arm
    of the file-level comparison.  kind None = the fastest arithmetic available (GFNI port, else reference C kernel)."""
    if kind is None:
        kind = 1 if gfni_level() else 0
    if kind == 0 and not ref_available():
        return -38
    return lib().orc_generate_ec_files_mt(base.encode(), kind, threads or os.cpu_count() or 1, large, small, k, m)
