"""oracle/rs_numpy.py — CPU ORACLE, second statement (test infrastructure, NOT product code).

An independent numpy restatement of the reference's RS arithmetic and shard layout, written
separately from rs_oracle.c so the two can check each other.  Only tests/, smoke() and bench.py's
cpu_baseline leg may import this.  Citations are relative to /root/reference; "rse/" abbreviates
seaweed-volume/vendor/reed-solomon-erasure/.
"""
from __future__ import annotations

import numpy as np

POLY = 29  # rse/build.rs:11 (0x11D without the x^8 term)


def _tables():
    log = np.zeros(256, dtype=np.int64)
    exp = np.zeros(510, dtype=np.int64)
    b = 1
    for lg in range(255):  # rse/build.rs:13-28
        log[b] = lg
        b <<= 1
        if b >= 256:
            b = (b - 256) ^ POLY
    for i in range(1, 256):  # rse/build.rs:30-42
        exp[log[i]] = i
        exp[log[i] + 255] = i
    a = np.arange(256)
    mul = exp[(log[a][:, None] + log[a][None, :])]  # rse/build.rs:44-68
    mul[0, :] = 0
    mul[:, 0] = 0
    return log.astype(np.uint8), exp.astype(np.uint8), mul.astype(np.uint8)


LOG, EXP, MUL = _tables()


def gf_mul(a: int, b: int) -> int:  # rse/src/galois_8.rs:67-69
    return int(MUL[a, b])


def gf_div(a: int, b: int) -> int:  # rse/src/galois_8.rs:72-86
    if a == 0:
        return 0
    if b == 0:
        raise ZeroDivisionError("Divisor is 0")
    return int(EXP[(int(LOG[a]) - int(LOG[b])) % 255])


def gf_exp(a: int, n: int) -> int:  # rse/src/galois_8.rs:89-103
    if n == 0:
        return 1
    if a == 0:
        return 0
    return int(EXP[(int(LOG[a]) * n) % 255])


def mat_mul(a: np.ndarray, b: np.ndarray) -> np.ndarray:  # rse/src/matrix.rs:119-139
    out = np.zeros((a.shape[0], b.shape[1]), dtype=np.uint8)
    for i in range(a.shape[1]):
        out ^= MUL[a[:, i][:, None], b[i, :][None, :]]
    return out


def mat_inv(m: np.ndarray) -> np.ndarray:  # rse/src/matrix.rs:195-261
    n = m.shape[0]
    w = np.concatenate([m.astype(np.uint8), np.eye(n, dtype=np.uint8)], axis=1)
    for r in range(n):
        if w[r, r] == 0:
            for rb in range(r + 1, n):
                if w[rb, r] != 0:
                    w[[r, rb]] = w[[rb, r]]
                    break
        if w[r, r] == 0:
            raise ValueError("singular matrix")
        if w[r, r] != 1:
            w[r] = MUL[gf_div(1, int(w[r, r])), w[r]]
        for rb in range(r + 1, n):
            if w[rb, r] != 0:
                w[rb] ^= MUL[int(w[rb, r]), w[r]]
    for d in range(n):
        for ra in range(d):
            if w[ra, d] != 0:
                w[ra] ^= MUL[int(w[ra, d]), w[d]]
    return w[:, n:].copy()


def vandermonde(rows: int, cols: int) -> np.ndarray:  # rse/src/matrix.rs:263-276
    return np.array([[gf_exp(r, c) for c in range(cols)] for r in range(rows)], dtype=np.uint8)


def build_matrix(k: int, total: int) -> np.ndarray:  # rse/src/core.rs:431-437
    v = vandermonde(total, k)
    return mat_mul(v, mat_inv(v[:k]))


def apply_rows(rows: np.ndarray, inputs: list[np.ndarray]) -> list[np.ndarray]:
    """out[p] = XOR_i rows[p,i] ⊗ inputs[i]   (rse/src/core.rs:484-512)."""
    outs = []
    for p in range(rows.shape[0]):
        acc = np.zeros_like(inputs[0])
        for i, x in enumerate(inputs):
            acc ^= MUL[int(rows[p, i])][x]
        outs.append(acc)
    return outs


def encode(k: int, m: int, data: list[np.ndarray]) -> list[np.ndarray]:  # rse/src/core.rs:600-635
    return apply_rows(build_matrix(k, k + m)[k:], data)


def reconstruct(k: int, m: int, shards: list[np.ndarray | None], data_only: bool = False):
    """rse/src/core.rs:736-926: first k present shards → inverse → missing data → missing parity."""
    total = k + m
    present = [s is not None for s in shards]
    if all(present):
        return list(shards)
    if sum(present) < k:
        raise ValueError("too few shards present")
    gen = build_matrix(k, total)
    valid = [i for i in range(total) if present[i]][:k]
    dec = mat_inv(gen[valid])
    sub = [shards[i] for i in valid]
    out = list(shards)
    missing_data = [j for j in range(k) if not present[j]]
    if missing_data:
        for j, s in zip(missing_data, apply_rows(dec[missing_data], sub)):
            out[j] = s
    if not data_only:
        missing_par = [p for p in range(k, total) if not present[p]]
        if missing_par:
            for p, s in zip(missing_par, apply_rows(gen[missing_par], out[:k])):
                out[p] = s
    return out


def fused_reconstruct_rows(k: int, m: int, present: list[bool], data_only: bool = False):
    """The single R×k matrix over the first k present shards that yields every missing shard:
    decode rows for missing data, parity_row·decode for missing parity."""
    total = k + m
    gen = build_matrix(k, total)
    valid = [i for i in range(total) if present[i]][:k]
    dec = mat_inv(gen[valid])
    missing = [i for i in range(total) if not present[i] and (i < k or not data_only)]
    rows = [dec[i] if i < k else mat_mul(gen[i : i + 1], dec)[0] for i in missing]
    return valid, missing, np.array(rows, dtype=np.uint8).reshape(len(missing), k)


def expected_shard_size(dat_size: int, k: int = 10, large: int = 1 << 30, small: int = 1 << 20) -> int:
    """weed/storage/disk_location_ec.go:428-448"""
    nlarge = dat_size // (large * k)
    size = nlarge * large
    rem = dat_size - nlarge * large * k
    if rem > 0:
        size += -(-rem // (small * k)) * small
    return size


def encode_dat_image(dat: np.ndarray, k: int = 10, m: int = 4, large: int = 1 << 30, small: int = 1 << 20):
    """weed/storage/erasure_coding/ec_encoder.go:280-321 on a memory image (batching is
    result-neutral: the code is column-wise)."""
    n = int(dat.shape[0])
    shard_size = expected_shard_size(n, k, large, small)
    data = [np.zeros(shard_size, dtype=np.uint8) for _ in range(k)]
    remaining, processed, written = n, 0, 0
    while remaining >= large * k:
        for i in range(k):
            data[i][written : written + large] = dat[processed + i * large : processed + (i + 1) * large]
        remaining -= large * k
        processed += large * k
        written += large
    while remaining > 0:
        for i in range(k):
            lo = processed + i * small
            chunk = dat[lo : min(lo + small, n)] if lo < n else dat[:0]
            data[i][written : written + len(chunk)] = chunk
        remaining -= small * k
        processed += small * k
        written += small
    return data + encode(k, m, data)


def locate_data(large: int, small: int, shard_dat_size: int, offset: int, size: int, k: int = 10):
    """weed/storage/erasure_coding/ec_locate.go:16-85 → list of
    (block_index, inner_block_offset, size, is_large_block, large_block_rows_count)."""
    nlarge_rows = shard_dat_size // large
    if offset < nlarge_rows * large * k:
        is_large, block_index, inner = True, offset // large, offset % large
    else:
        off = offset - nlarge_rows * large * k
        is_large, block_index, inner = False, off // small, off % small
    out = []

    def advance(bi, il):
        bi += 1
        if il and bi == nlarge_rows * k:
            return 0, False
        return bi, il

    while size > 0:
        remaining = (large if is_large else small) - inner
        if remaining <= 0:
            block_index, is_large = advance(block_index, is_large)
            inner = 0
            continue
        if size <= remaining:
            out.append((block_index, inner, size, is_large, nlarge_rows))
            return out
        out.append((block_index, inner, remaining, is_large, nlarge_rows))
        size -= remaining
        block_index, is_large = advance(block_index, is_large)
        inner = 0
    return out


def interval_to_shard(iv, large: int, small: int, k: int = 10):
    """weed/storage/erasure_coding/ec_locate.go:87-98 → (shard_id, offset in shard file)."""
    block_index, inner, _size, is_large, nlarge_rows = iv
    row = block_index // k
    off = inner + (row * large if is_large else nlarge_rows * large + row * small)
    return block_index % k, off


def synth(byte_offset: int, n: int, seed: int) -> np.ndarray:
    """splitmix64 over the 8-byte word index, little-endian bytes."""
    first, last = byte_offset // 8, (byte_offset + n + 7) // 8
    with np.errstate(over="ignore"):
        j = np.arange(first, last, dtype=np.uint64)
        z = np.uint64(seed) + (j + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    b = z.astype("<u8").view(np.uint8)
    lo = byte_offset - first * 8
    return b[lo : lo + n].copy()


# ---- index files either side of the path (no GF math) ------------------------------------------------
# Entry = 8-byte id, 4-byte offset (units of 8 B), 4-byte size, big-endian
# (weed/storage/types/needle_types.go:58-64, offset_4bytes.go:14-60).

TOMBSTONE = -1


def _entries(raw: bytes):
    for off in range(0, len(raw) - len(raw) % 16, 16):
        key = int.from_bytes(raw[off:off + 8], "big")
        offset = int.from_bytes(raw[off + 8:off + 12], "big")
        size = int.from_bytes(raw[off + 12:off + 16], "big", signed=True)
        yield key, offset, size


def _entry(key: int, offset: int, size: int) -> bytes:
    return key.to_bytes(8, "big") + offset.to_bytes(4, "big") + (size & 0xFFFFFFFF).to_bytes(4, "big")


def sorted_ecx_from_idx(idx: bytes) -> bytes:
    """readNeedleMap + AscendingVisit (weed/storage/erasure_coding/ec_encoder.go:31-58,379-396)."""
    live = {}
    for key, offset, size in _entries(idx):
        if offset != 0 and not (size < 0):
            live[key] = (offset, size)
        else:
            live.pop(key, None)
    return b"".join(_entry(k, *live[k]) for k in sorted(live))


def fold_ecj_into_ecx(ecx: bytes, ecj: bytes) -> bytes:
    """RebuildEcxFile (weed/storage/erasure_coding/ec_volume_delete.go:95-142): binary search, tombstone size."""
    out = bytearray(ecx)
    keys = [k for k, _, _ in _entries(ecx)]
    for off in range(0, len(ecj) - len(ecj) % 8, 8):
        nid = int.from_bytes(ecj[off:off + 8], "big")
        lo, hi = 0, len(keys)
        while lo < hi:
            mid = (lo + hi) // 2
            if keys[mid] == nid:
                out[mid * 16 + 12:mid * 16 + 16] = (TOMBSTONE & 0xFFFFFFFF).to_bytes(4, "big")
                break
            if keys[mid] < nid:
                lo = mid + 1
            else:
                hi = mid
    return bytes(out)


def idx_from_ec_index(ecx: bytes, ecj: bytes) -> bytes:
    """WriteIdxFileFromEcIndex (weed/storage/erasure_coding/ec_decoder.go:35-60)."""
    out = bytearray(ecx)
    for off in range(0, len(ecj) - len(ecj) % 8, 8):
        out += _entry(int.from_bytes(ecj[off:off + 8], "big"), 0, TOMBSTONE)
    return bytes(out)


def find_dat_file_size(ecx: bytes, version: int) -> int:
    """FindDatFileSize (ec_decoder.go:65-92) with GetActualSize (needle/needle_read.go:292-294,
    needle_read_tail.go:36-50): header 16 + size + checksum 4 (+ timestamp 8 for v3) + padding 1..8."""
    best = 8  # SuperBlockSize
    for _key, offset, size in _entries(ecx):
        if size < 0:
            continue
        fixed = 16 + size + 4 + (8 if version == 3 else 0)
        best = max(best, offset * 8 + fixed + (8 - fixed % 8))
    return best


def _i32(v: int) -> int:
    v &= 0xFFFFFFFF
    return v - (1 << 32) if v & 0x80000000 else v


def get_actual_size(size: int, version: int) -> int:
    """needle.GetActualSize with the reference's types (needle_read.go:292-294, needle_read_tail.go:36-49):
    PaddingLength is evaluated in Size (int32) arithmetic — it wraps, and Go's % keeps the dividend's sign —
    while NeedleBodyLength adds in int64."""
    tail = 12 if version == 3 else 4
    s = _i32(16 + size + tail)
    rem = abs(s) % 8
    rem = -rem if s < 0 else rem
    return 16 + size + tail + (8 - rem)


def check_index_file(raw: bytes, version: int = 3):
    """idx.CheckIndexFile (weed/storage/idx/check.go:36-111): entries in (offset, size) order; a finding for every
    neighbour that starts at or before the end of its predecessor; a finding when the file is not a whole number of
    entries.  Returns (entries processed, findings)."""
    es = [(i, k, o * 8, s) for i, (k, o, s) in enumerate(_entries(raw))]
    es.sort(key=lambda e: (e[2], e[3]))                      # stable, like the run of equal keys in the fixtures
    errs = []
    for j in range(1, len(es)):
        idx, key, off, size = es[j]
        _, lkey, loff, lsize = es[j - 1]
        end = off + (get_actual_size(size, version) - 1 if get_actual_size(size, version) else 0)
        lend = loff + (get_actual_size(lsize, version) - 1 if get_actual_size(lsize, version) else 0)
        if off <= lend:
            errs.append(f"needle {key} (#{idx + 1}) at [{off}-{end}] overlaps needle {lkey} at [{loff}-{lend}]")
    if len(es) * 16 != len(raw):
        errs.append(f"expected an index file of size {len(raw)}, got {len(es) * 16}")
    return len(es), errs


def read_needle_record(dat: np.ndarray, offset: int, size: int, version: int = 3) -> np.ndarray:
    """The bytes Store.ReadEcShardNeedle hands to the needle parser (store_ec.go:252-290): LocateEcShardNeedle
    passes GetActualSize(size) to LocateEcShardNeedleInterval, which applies GetActualSize again
    (ec_volume.go:395,414), so the read runs past the record; past the end of the volume the shards hold the
    zero padding of the last small row."""
    want = get_actual_size(get_actual_size(size, version), version)
    padded = np.concatenate([dat, np.zeros(want + 16, dtype=np.uint8)])
    return padded[offset:offset + want]
