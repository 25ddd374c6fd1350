"""The reference's own test suites for this path, re-run against the B200 engine through the Go-named
mirror (seaweedfs_b200.erasure_coding) — same sizes, same offsets, same assertions:

  TestEcReadRoundTrip, TestEcOffByOneBug_Issue8947, TestEcDecodeDatRoundTrip
                                                   weed/storage/erasure_coding/ec_roundtrip_test.go:22-345
  TestEcConsistency_WritesBetweenEncodeAndEcx, TestEcConsistency_ExactLargeRowEncoding
                                                   weed/storage/erasure_coding/ec_consistency_test.go:31-175
  TestRecoverOneRemoteEcShardInterval_{SufficientShards, InsufficientShards, ReconstructDataSlicing,
  ParityShardRecovery}                             weed/storage/store_ec_recovery_test.go:25-82,194-300

The Go tests only assert round trips; here every produced shard is additionally compared with the oracle."""
import os

import numpy as np
import pytest

from oracle import rs_numpy as rn

pytestmark = pytest.mark.gpu

LARGE, SMALL = 10000, 100            # largeBlockSize / smallBlockSize of ec_test.go:19-22
K = 10
LARGE_ROW, SMALL_ROW = LARGE * K, SMALL * K


def encode_files(ec, oracle, tmp_path, dat, name, large=LARGE, small=SMALL, buffer=SMALL):
    base = str(tmp_path / name)
    dat.tofile(base + ".dat")
    ec.generateEcFiles(base, buffer, large, small)
    shards = [np.fromfile(base + ec.ToExt(i), dtype=np.uint8) for i in range(14)]
    want = oracle.encode_dat_image(dat, buffer_size=buffer, large=large, small=small)
    for i in range(14):
        assert shards[i].shape == want[i].shape and (shards[i] == want[i]).all(), f"{name}: shard {i} differs from the oracle"
    return base, shards


def assemble(ec, shards, intervals, large=LARGE, small=SMALL):
    """assembleFromIntervals (ec_roundtrip_test.go:372-388): None on a short read"""
    out = []
    for iv in intervals:
        sid, soff = ec.interval_to_shard(iv, large, small)
        piece = shards[sid][soff:soff + iv[2]]
        if len(piece) != iv[2]:
            return None
        out.append(piece)
    return np.concatenate(out) if out else np.zeros(0, dtype=np.uint8)


def collect_test_offsets(dat_size, read_size, boundary, large, small):      # ec_roundtrip_test.go:346-369
    offs = [0]
    if dat_size > read_size:
        offs.append(dat_size // 2)
    if 0 < boundary < dat_size:
        for delta in (-large, -small, -1, 0, 1, small, large):
            off = boundary + delta
            if off >= 0 and off + read_size <= dat_size:
                offs.append(off)
    if dat_size > read_size:
        offs.append(dat_size - read_size)
    return offs


ROUND_TRIP_SIZES = {
    "1_large_row_exact": LARGE_ROW, "2_large_rows_exact": 2 * LARGE_ROW, "3_large_rows_exact": 3 * LARGE_ROW,
    "1_large_row_plus_1": LARGE_ROW + 1, "2_large_rows_plus_small": 2 * LARGE_ROW + SMALL_ROW,
    "1_large_row_plus_half_small": LARGE_ROW + SMALL_ROW // 2,
    "just_under_1_large_row": LARGE_ROW - 1, "just_under_2_large_rows": 2 * LARGE_ROW - 1,
    "small_only": SMALL_ROW * 3, "small_single_row": SMALL_ROW,
    "boundary_spanning": LARGE_ROW + SMALL_ROW * 5 + 50,
}


@pytest.mark.parametrize("name", list(ROUND_TRIP_SIZES))
def test_ec_read_round_trip(cuda, swec, oracle, tmp_path, name):
    """TestEcReadRoundTrip / testEcRead"""
    ec = swec.erasure_coding
    dat_size = ROUND_TRIP_SIZES[name]
    dat = np.random.default_rng(dat_size).integers(0, 256, dat_size, dtype=np.uint8)
    _, shards = encode_files(ec, oracle, tmp_path, dat, "rt")
    shard_dat_size = dat_size // K
    boundary = (dat_size // LARGE_ROW) * LARGE_ROW
    read_size = SMALL // 2
    for off in collect_test_offsets(dat_size, read_size, boundary, LARGE, SMALL):
        got = assemble(ec, shards, ec.LocateData(LARGE, SMALL, shard_dat_size, off, read_size))
        assert got is not None and (got == dat[off:off + read_size]).all(), (name, off)
        # the ecdFileSize-1 fallback of old volumes: allowed to miss only on exact multiples (the Go test logs it)
        fb = assemble(ec, shards, ec.LocateData(LARGE, SMALL, len(shards[0]) - 1, off, read_size))
        if dat_size % LARGE_ROW:
            assert fb is not None and (fb == dat[off:off + read_size]).all(), (name, off, "fallback")


def test_ec_off_by_one_bug_issue_8947(cuda, swec, oracle, tmp_path):
    """TestEcOffByOneBug_Issue8947: exactly two large rows; the fixed row count reads the 2nd row correctly, the
    old (shardDatSize-1)/large count misclassifies it as small blocks."""
    ec = swec.erasure_coding
    dat_size = 2 * LARGE_ROW
    dat = np.random.default_rng(8947).integers(0, 256, dat_size, dtype=np.uint8)
    _, shards = encode_files(ec, oracle, tmp_path, dat, "bug")
    shard_dat_size = dat_size // K
    assert shard_dat_size // LARGE == 2 and (shard_dat_size - 1) // LARGE == 1
    off, read_size = LARGE_ROW + LARGE + 50, SMALL // 2
    fixed = ec.LocateData(LARGE, SMALL, shard_dat_size, off, read_size)
    assert fixed[0][3] is True                                            # IsLargeBlock
    assert (assemble(ec, shards, fixed) == dat[off:off + read_size]).all()
    buggy = ec.LocateData(LARGE, SMALL, shard_dat_size - 1, off, read_size)   # what the old formula computed
    assert buggy[0][3] is False
    wrong = assemble(ec, shards, buggy)
    assert wrong is None or not (wrong == dat[off:off + read_size]).all()


@pytest.mark.parametrize("dat_size", [1000, 10 * (1 << 20), 10 * (1 << 20) + 500])
def test_ec_decode_dat_round_trip(cuda, swec, oracle, tmp_path, dat_size):
    """TestEcDecodeDatRoundTrip: production block sizes, WriteDatFile gives the .dat back."""
    ec = swec.erasure_coding
    dat = np.random.default_rng(dat_size).integers(0, 256, dat_size, dtype=np.uint8)
    base, _ = encode_files(ec, oracle, tmp_path, dat, "dec", large=1 << 30, small=1 << 20, buffer=256 * 1024)
    ec.WriteDatFile(base + "_decoded", dat_size, [base + ec.ToExt(i) for i in range(10)])
    back = np.fromfile(base + "_decoded.dat", dtype=np.uint8)
    assert back.shape == dat.shape and (back == dat).all()


def test_ec_consistency_exact_large_row_encoding(cuda, swec, oracle, tmp_path):
    """TestEcConsistency_ExactLargeRowEncoding: one large row exactly ⇒ every shard is largeBlockSize long and
    every smallBlockSize chunk reads back through LocateData."""
    ec = swec.erasure_coding
    dat = np.random.default_rng(77).integers(0, 256, LARGE_ROW, dtype=np.uint8)
    _, shards = encode_files(ec, oracle, tmp_path, dat, "exact")
    assert all(len(s) == LARGE for s in shards)
    for off in range(0, LARGE_ROW, SMALL):
        got = assemble(ec, shards, ec.LocateData(LARGE, SMALL, LARGE_ROW // K, off, SMALL))
        assert got is not None and (got == dat[off:off + SMALL]).all(), off


def test_ec_consistency_writes_between_encode_and_ecx(cuda, swec, oracle, tmp_path):
    """TestEcConsistency_WritesBetweenEncodeAndEcx: a needle appended to .dat after the shards were generated is
    in a late .ecx but not in the shards (why .ecx is written FIRST); original data still reads."""
    ec = swec.erasure_coding
    dat_size = LARGE_ROW + SMALL_ROW * 3
    rng = np.random.default_rng(5)
    dat = rng.integers(0, 256, dat_size, dtype=np.uint8)
    base, shards = encode_files(ec, oracle, tmp_path, dat, "consistency")
    extra = rng.integers(1, 256, 5000, dtype=np.uint8)
    with open(base + ".dat", "ab") as f:
        f.write(extra.tobytes())
    open(base + ".idx", "wb").write(rn._entry(1, 0, dat_size) + rn._entry(2, dat_size // 8, len(extra)))
    ec.WriteSortedFileFromIdx(base, ".ecx")
    fixed = 16 + len(extra) + 4 + 8
    actual = fixed + (8 - fixed % 8)
    got = assemble(ec, shards, ec.LocateData(LARGE, SMALL, len(shards[0]) - 1, dat_size, actual))
    assert got is None or not (got[:len(extra)] == extra).all()          # the needle is NOT in the shards
    first = assemble(ec, shards, ec.LocateData(LARGE, SMALL, dat_size // K, 0, SMALL))
    assert (first == dat[:SMALL]).all()


# ---- store_ec_recovery_test.go -------------------------------------------------------------------

def _recovery_shards(oracle, n=1024):
    data = [np.array([(i + j) & 255 for j in range(n)], dtype=np.uint8) for i in range(10)]   # byte(i + j), :208-213
    return data + oracle.encode(10, 4, data)


@pytest.mark.parametrize("available", [list(range(10)), list(range(1, 11)), [0, 1, 2, 3, 4, 5, 10, 11, 12, 13],
                                       list(range(4, 14))])
def test_recover_with_sufficient_shards(cuda, swec, oracle, available):
    """TestRecoverOneRemoteEcShardInterval_SufficientShards / _ReconstructDataSlicing: any 10 of 14 recover the
    missing data shards through ReconstructData; the recovered bytes equal the originals."""
    enc = swec.erasure_coding.Encoder(10, 4, device=0)
    full = _recovery_shards(oracle)
    bufs = [full[i].copy() if i in available else None for i in range(14)]
    enc.ReconstructData(bufs)
    for i in range(10):
        assert (bufs[i] == full[i]).all(), i
    enc.close()


def test_recover_with_insufficient_shards(cuda, swec, oracle):
    """TestRecoverOneRemoteEcShardInterval_InsufficientShards: 9 shards ⇒ ErrTooFewShards, nothing allocated."""
    enc = swec.erasure_coding.Encoder(10, 4, device=0)
    full = _recovery_shards(oracle)
    bufs = [full[i].copy() if i < 9 else None for i in range(14)]
    with pytest.raises(swec.SwecError) as e:
        enc.ReconstructData(bufs)
    assert e.value.name == "SWEC_ERR_TOO_FEW_SHARDS"
    assert all(b is None for b in bufs[9:])
    enc.close()


@pytest.mark.parametrize("parity_shard", [10, 11, 12, 13])
def test_parity_shard_recovery(cuda, swec, oracle, parity_shard):
    """TestRecoverOneRemoteEcShardInterval_ParityShardRecovery (inputs byte(i * j)): Reconstruct rebuilds a lost
    parity shard from the ten data shards."""
    n = 512
    data = [np.array([(i * j) & 255 for j in range(n)], dtype=np.uint8) for i in range(10)]
    full = data + oracle.encode(10, 4, data)
    enc = swec.erasure_coding.Encoder(10, 4, device=0)
    bufs = [None if i == parity_shard else full[i].copy() for i in range(14)]
    enc.Reconstruct(bufs)
    assert bufs[parity_shard] is not None and (bufs[parity_shard] == full[parity_shard]).all()
    enc.close()


# ---- disk_location_ec_realworld_test.go ------------------------------------------------------------

@pytest.mark.parametrize("dat_size", [1, 1024, 10 * 1024, 1 << 20, (1 << 20) + 1, 9 * (1 << 20) + 900 * 1024,
                                      10 * (1 << 20) + 100 * 1024,
                                      # TestCalculateExpectedShardSizeWithRealEncoding (:13-129)
                                      5 << 20, 10 << 20, 15 << 20, 50 << 20, 100 << 20, 512 << 20])
def test_calculate_expected_shard_size_with_real_encoding(cuda, swec, oracle, tmp_path, dat_size):
    """TestCalculateExpectedShardSizeEdgeCases / …WithRealEncoding (disk_location_ec_realworld_test.go:13-200): WriteEcFiles on
    byte(i % 256) data, every shard file has the size calculateExpectedShardSize predicts."""
    ec = swec.erasure_coding
    dat = np.tile(np.arange(256, dtype=np.uint8), dat_size // 256 + 1)[:dat_size]        # byte(i % 256)
    base, shards = encode_files(ec, oracle, tmp_path, dat, "edge", large=1 << 30, small=1 << 20, buffer=256 * 1024)
    want = ec.expected_shard_size(dat_size)
    assert all(os.path.getsize(base + ec.ToExt(i)) == want for i in range(14))


@pytest.mark.parametrize("large,small,dat_size", [(32 << 20, 16 << 20, (330 << 20) + 12345),     # small block > the 8 MiB slot
                                                  (1 << 20, 1 << 20, (25 << 20) + 1),             # large == small
                                                  (3 << 20, 1 << 20, 10 * (3 << 20) + 10 * (1 << 20) * 9 + 5)])  # 1 large + 10 small rows
def test_generate_ec_files_unusual_block_sizes(cuda, swec, oracle, tmp_path, large, small, dat_size):
    """generateEcFiles takes its block sizes as arguments (ec_encoder.go:110): the row batching of the file
    pipeline must hold for small blocks larger than a staging slot, equal block sizes and more small rows than
    fit one slot."""
    ec = swec.erasure_coding
    dat = np.random.default_rng(dat_size).integers(0, 256, dat_size, dtype=np.uint8)
    base, shards = encode_files(ec, oracle, tmp_path, dat, "blk", large=large, small=small, buffer=1 << 20)
    assert len(shards[0]) == ec.expected_shard_size(dat_size, 10, large, small)
    for i in (2, 11):
        os.remove(base + ec.ToExt(i))
    assert ec.rebuild_ec_files(base) == [2, 11]
    want = oracle.encode_dat_image(dat, buffer_size=1 << 20, large=large, small=small)
    for i in (2, 11):
        assert (np.fromfile(base + ec.ToExt(i), dtype=np.uint8) == want[i]).all()
