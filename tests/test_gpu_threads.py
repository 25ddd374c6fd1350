"""Threading contract of the C ABI (include/swec.h): concurrent callers on separate encoder handles
and on one shared handle, like Go goroutines running several volumes at once (weed/shell/common.go:11)."""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_concurrent_handles_and_shared_handle(cuda, swec, oracle):
    ec = swec.erasure_coding
    rng = np.random.default_rng(77)
    n = 700_001
    cases = []
    for _ in range(6):
        data = [rng.integers(0, 256, n, dtype=np.uint8) for _ in range(10)]
        cases.append((data, oracle.encode(10, 4, data)))
    shared = ec.Encoder(10, 4, device=0)
    errors = []

    def worker(idx, enc):
        try:
            data, want = cases[idx]
            for _ in range(3):
                shards = [d.copy() for d in data] + [np.zeros(n, dtype=np.uint8) for _ in range(4)]
                enc.encode(shards)
                assert all((a == b).all() for a, b in zip(shards[10:], want))
                holes = list(shards)
                holes[idx % 10] = None
                holes[10 + idx % 4] = None
                enc.reconstruct(holes)
                assert all((a == b).all() for a, b in zip(holes, shards))
        except Exception as ex:  # noqa: BLE001
            errors.append((idx, repr(ex)))

    threads = [threading.Thread(target=worker, args=(i, ec.Encoder(10, 4, device=0) if i % 2 else shared))
               for i in range(6)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors


def test_one_process_many_gpus(cuda, swec, oracle):
    """A volume server is ONE process driving every GPU of the box: volume v goes to GPU v mod N, each on
    its own handle and OS thread (no collective).  Needs >= 2 GPUs."""
    torch = cuda
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("single-GPU box")
    ec = swec.erasure_coding
    rng = np.random.default_rng(5)
    n = 3_000_000 + 5
    volumes = []
    for v in range(2 * ngpu):
        data = [rng.integers(0, 256, n, dtype=np.uint8) for _ in range(10)]
        volumes.append((data, oracle.encode(10, 4, data)))
    errors = []

    def worker(v):
        try:
            enc = ec.Encoder(10, 4, device=v % ngpu)
            data, want = volumes[v]
            shards = [d.copy() for d in data] + [np.zeros(n, dtype=np.uint8) for _ in range(4)]
            enc.encode(shards)
            assert all((a == b).all() for a, b in zip(shards[10:], want))
            holes = list(shards)
            for i in (0, 1, 2, 3):
                holes[i] = None
            enc.reconstruct(holes)
            assert all((a == b).all() for a, b in zip(holes, shards))
            # device-resident on that GPU too
            with torch.cuda.device(v % ngpu):
                d = [torch.from_numpy(x).cuda() for x in data]
                p = [torch.zeros(n, dtype=torch.uint8, device="cuda") for _ in range(4)]
                enc.encode_device([t.data_ptr() for t in d], [t.data_ptr() for t in p], n,
                                  torch.cuda.current_stream().cuda_stream)
                torch.cuda.synchronize()
                assert all((a.cpu().numpy() == b).all() for a, b in zip(p, want))
            enc.close()
        except Exception as ex:  # noqa: BLE001
            errors.append((v, repr(ex)))

    threads = [threading.Thread(target=worker, args=(v,)) for v in range(len(volumes))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
