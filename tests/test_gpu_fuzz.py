"""Property-based parity (hypothesis, derandomised): the host API of the C ABI against the CPU oracle over
randomly drawn codes (any k + m <= 32, as .vif allows — ec_encoder.go:77-91), shard lengths (including
< 16-byte and odd tails), unaligned caller buffers, erasure patterns and ReconstructData vs Reconstruct —
the combinations the hand-written cases of test_gpu_parity.py do not enumerate."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

pytestmark = pytest.mark.gpu

_encoders = {}


def _encoder(swec, k, m):
    if (k, m) not in _encoders:
        _encoders[(k, m)] = swec.erasure_coding.Encoder(k, m, device=0)
    return _encoders[(k, m)]


codes = st.one_of(st.just((10, 4)), st.tuples(st.integers(1, 20), st.integers(1, 12)).filter(lambda c: c[0] + c[1] <= 32))
lengths = st.one_of(st.integers(1, 300), st.integers(4000, 70000), st.sampled_from([16, 4096, 262144, 262144 + 1]))
COMMON = dict(deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))


def _shards(rng, k, n, shift):
    """k data arrays of n bytes living at an arbitrary offset inside larger buffers (unaligned callers)."""
    out = []
    for _ in range(k):
        backing = rng.integers(0, 256, n + 32, dtype=np.uint8)
        out.append(backing[shift:shift + n])
    return out


@settings(max_examples=60, **COMMON)
@given(code=codes, n=lengths, shift=st.integers(0, 15), seed=st.integers(0, 2**31))
def test_encode_matches_oracle(cuda, swec, oracle, code, n, shift, seed):
    k, m = code
    rng = np.random.default_rng(seed)
    data = _shards(rng, k, n, shift)
    want = oracle.encode(k, m, [np.ascontiguousarray(d) for d in data])
    before = [d.copy() for d in data]
    parity = [np.full(n + 32, 0xA5, dtype=np.uint8)[shift:shift + n] for _ in range(m)]
    _encoder(swec, k, m).encode(data + parity)
    for d, b in zip(data, before):
        assert (d == b).all(), "Encode must not touch the data shards"
    for p in range(m):
        assert (parity[p] == want[p]).all(), (k, m, n, shift, p)
        assert (parity[p].base[:shift] == 0xA5).all() and (parity[p].base[shift + n:] == 0xA5).all()


@settings(max_examples=60, **COMMON)
@given(code=codes, n=lengths, data_only=st.booleans(), seed=st.integers(0, 2**31), data=st.data())
def test_reconstruct_matches_oracle(cuda, swec, oracle, code, n, data_only, seed, data):
    k, m = code
    rng = np.random.default_rng(seed)
    full = [rng.integers(0, 256, n, dtype=np.uint8) for _ in range(k)]
    full += oracle.encode(k, m, full)
    n_lost = data.draw(st.integers(1, m))
    lost = data.draw(st.lists(st.integers(0, k + m - 1), min_size=n_lost, max_size=n_lost, unique=True))
    shards = [None if i in lost else full[i].copy() for i in range(k + m)]
    _encoder(swec, k, m).reconstruct(shards, data_only=data_only)
    for i in range(k + m):
        if i in lost and i >= k and data_only:
            continue                                   # ReconstructData fills data shards only
        assert shards[i] is not None and (shards[i] == full[i]).all(), (k, m, n, sorted(lost), i)


@settings(max_examples=25, **COMMON)
@given(code=codes, n=st.integers(1, 5000), seed=st.integers(0, 2**31), data=st.data())
def test_verify_detects_any_single_byte_corruption(cuda, swec, oracle, code, n, seed, data):
    k, m = code
    rng = np.random.default_rng(seed)
    full = [rng.integers(0, 256, n, dtype=np.uint8) for _ in range(k)]
    full += oracle.encode(k, m, full)
    enc = _encoder(swec, k, m)
    assert enc.verify(full)
    victim, pos = data.draw(st.integers(0, k + m - 1)), data.draw(st.integers(0, n - 1))
    full[victim][pos] ^= data.draw(st.integers(1, 255))
    assert not enc.verify(full)


@settings(max_examples=40, **COMMON)
@given(k=st.integers(1, 14), small_pow=st.integers(0, 6), mult=st.integers(2, 40), size=st.integers(0, 200000),
       seed=st.integers(0, 2**31))
def test_volume_device_layout_matches_oracle(cuda, swec, oracle, k, small_pow, mult, size, seed):
    """encodeDatFile's two-tier row walk for arbitrary block sizes and volume sizes (ec_encoder.go:280-321)."""
    torch = cuda
    m = 3
    small = 16 << small_pow
    large = small * mult
    rng = np.random.default_rng(seed)
    dat = rng.integers(0, 256, size, dtype=np.uint8)
    want = oracle.encode_dat_image(dat, k=k, m=m, buffer_size=16, large=large, small=small)
    ec = swec.erasure_coding
    shard = ec.expected_shard_size(size, k, large, small)
    assert shard == len(want[0])
    d = torch.from_numpy(np.concatenate([dat, np.zeros(16, dtype=np.uint8)])).cuda()
    par = [torch.full((max(shard, 1),), 0x5A, dtype=torch.uint8, device="cuda") for _ in range(m)]
    enc = _encoder(swec, k, m)
    enc.encode_volume_device(d.data_ptr(), size, [p.data_ptr() for p in par], torch.cuda.current_stream().cuda_stream,
                             large_block=large, small_block=small)
    torch.cuda.synchronize()
    for p in range(m):
        assert (par[p][:shard].cpu().numpy() == want[k + p]).all(), (k, large, small, size, p)
    one = torch.empty(max(shard, 1), dtype=torch.uint8, device="cuda")
    sid = seed % k
    enc.extract_data_shard_device(d.data_ptr(), size, sid, one.data_ptr(), torch.cuda.current_stream().cuda_stream,
                                  large_block=large, small_block=small)
    torch.cuda.synchronize()
    assert (one[:shard].cpu().numpy() == want[sid]).all()
