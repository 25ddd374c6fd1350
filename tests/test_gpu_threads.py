"""Threading contract of the C ABI (include/swec.h): concurrent callers on separate encoder handles
and on one shared handle, like Go goroutines running several volumes at once (weed/shell/common.go:11)."""
import os
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


@pytest.mark.parametrize("n", [1, 4095, 4096 * 3 + 17, 9_000_000 + 3])
def test_one_call_split_over_several_handles(cuda, swec, oracle, n):
    """swec_encode_multi / swec_reconstruct_multi cut the byte-column range across encoder handles — one per
    GPU when the box has several; on a single-GPU box three handles on device 0 exercise the same split
    (ranges of 4 KiB units, empty ranges when n is tiny).  Result = the single-handle call, bit for bit."""
    torch = cuda
    ec = swec.erasure_coding
    ngpu = torch.cuda.device_count()
    devices = list(range(ngpu)) if ngpu >= 2 else [0, 0, 0]
    grp = ec.EncoderGroup(10, 4, devices)
    rng = np.random.default_rng(n)
    data = [rng.integers(0, 256, n, dtype=np.uint8) for _ in range(10)]
    want = oracle.encode(10, 4, data)
    shards = [d.copy() for d in data] + [np.full(n, 0x77, dtype=np.uint8) for _ in range(4)]
    grp.encode(shards)
    assert all((a == b).all() for a, b in zip(shards[:10], data))
    assert all((a == b).all() for a, b in zip(shards[10:], want))
    holes = [s.copy() for s in shards]
    for i in (1, 6, 10, 12):
        holes[i] = None
    grp.reconstruct(holes)
    assert all((a == b).all() for a, b in zip(holes, shards))
    holes = [s.copy() for s in shards]
    holes[4] = holes[13] = None
    grp.reconstruct(holes, data_only=True)
    assert (holes[4] == shards[4]).all() and holes[13] is None
    grp.close()


def test_group_shard_buffers_are_usable_by_the_split_call(cuda, swec, oracle):
    """swec_alloc_pinned_shards: one pinned allocation whose column ranges sit near the GPU that takes them (NUMA
    binding is best effort); the column-split encode on those buffers equals the oracle, and they are freed whole."""
    import ctypes as C
    torch = cuda
    ec = swec.erasure_coding
    L = swec.lib()
    ngpu = torch.cuda.device_count()
    grp = ec.EncoderGroup(10, 4, list(range(ngpu)) if ngpu >= 2 else [0, 0])
    n = 5_000_000 + 77
    ptrs = (C.c_void_p * 14)()
    assert L.swec_alloc_pinned_shards(grp._arr, len(grp.encoders), 14, n, ptrs) == 0
    shards = [np.ctypeslib.as_array(C.cast(ptrs[i], C.POINTER(C.c_uint8)), shape=(n,)) for i in range(14)]
    assert all(ptrs[i] % 4096 == 0 for i in range(14))
    rng = np.random.default_rng(12)
    for s in shards[:10]:
        s[:] = rng.integers(0, 256, n, dtype=np.uint8)
    want = oracle.encode(10, 4, [s.copy() for s in shards[:10]])
    assert L.swec_encode_multi(grp._arr, len(grp.encoders), ptrs, n) == 0
    for p in range(4):
        assert (shards[10 + p] == want[p]).all()
    del shards
    L.swec_free_pinned(ptrs[0])
    grp.close()


def test_concurrent_file_level_calls_share_parked_rings(cuda, swec, oracle, tmp_path):
    """Several volumes encoded and rebuilt at once from different OS threads (the shell's ec.encode runs up to ten,
    weed/shell/common.go:11), twice over, so that later calls pick up staging rings parked by earlier ones while
    others are still running.  Every shard file equals the oracle's."""
    ec = swec.erasure_coding
    rng = np.random.default_rng(9)
    vols = []
    for v in range(4):
        size = int(rng.integers(3, 12)) * (1 << 20) + int(rng.integers(0, 4096))
        dat = rng.integers(0, 256, size, dtype=np.uint8)
        base = str(tmp_path / f"v{v}")
        dat.tofile(base + ".dat")
        vols.append((base, oracle.encode_dat_image(dat)))
    errors = []

    def worker(v):
        try:
            base, want = vols[v]
            for _ in range(2):
                ec.write_ec_files(base)
                for i in range(14):
                    assert (np.fromfile(base + ec.ToExt(i), dtype=np.uint8) == want[i]).all(), (v, i)
                lost = [(v + j * 3) % 14 for j in range(1 + v % 4)]
                for i in set(lost):
                    os.remove(base + ec.ToExt(i))
                assert ec.rebuild_ec_files(base) == sorted(set(lost))
                for i in range(14):
                    assert (np.fromfile(base + ec.ToExt(i), dtype=np.uint8) == want[i]).all(), (v, i, "rebuilt")
                ok, bad = ec.verify_ec_files(base)
                assert ok and not any(bad)
        except Exception as ex:  # noqa: BLE001
            errors.append((v, repr(ex)))

    threads = [threading.Thread(target=worker, args=(v,)) for v in range(len(vols))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors


def test_file_pipeline_write_error_surfaces_and_does_not_hang(cuda, swec, tmp_path):
    """A write that fails in the middle of a volume (disk full / file-size limit) must come back as SWEC_ERR_IO with
    the errno text — from whichever I/O thread hit it — and leave the pipeline shut down, not hung; the next call on
    the same device works (the ring of a failed run is not parked).  Run in a child process: RLIMIT_FSIZE."""
    import subprocess
    import sys
    import textwrap
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    child = textwrap.dedent(f"""
        import os, resource, signal, sys
        sys.path.insert(0, {root!r})
        import numpy as np
        import seaweedfs_b200
        from seaweedfs_b200 import erasure_coding as ec
        base = os.path.join({str(tmp_path)!r}, "5")
        np.random.default_rng(0).integers(0, 256, 60 << 20, dtype=np.uint8).tofile(base + ".dat")
        signal.signal(signal.SIGXFSZ, signal.SIG_IGN)             # EFBIG instead of a fatal signal
        resource.setrlimit(resource.RLIMIT_FSIZE, (3 << 20, resource.RLIM_INFINITY))
        try:
            ec.write_ec_files(base)
            print("NO ERROR")
        except seaweedfs_b200.SwecError as e:
            print("ERR", e.name, str(e))
        resource.setrlimit(resource.RLIMIT_FSIZE, (resource.RLIM_INFINITY, resource.RLIM_INFINITY))
        for i in range(14):
            os.remove(base + ec.ToExt(i))
        ec.write_ec_files(base)                                   # the engine is healthy afterwards
        print("RETRY OK", os.path.getsize(base + ".ec13"))
    """)
    r = subprocess.run([sys.executable, "-c", child], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert "ERR SWEC_ERR_IO" in r.stdout and "pwrite" in r.stdout, r.stdout[-2000:]
    assert "RETRY OK 6291456" in r.stdout, r.stdout[-2000:]


def test_device_spread_order_is_a_permutation_interleaved_over_numa_nodes(cuda, swec):
    """swec_device_spread_order: every device exactly once; consecutive entries alternate between the host's NUMA
    nodes for as long as both have devices left (what bench.py and swecPickDevice use to place n concurrent volumes)."""
    import ctypes as C
    torch = cuda
    n = torch.cuda.device_count()
    order, cnt = (C.c_int * 64)(), C.c_int(0)
    assert swec.lib().swec_device_spread_order(order, 64, C.byref(cnt)) == 0
    got = list(order[:cnt.value])
    assert sorted(got) == list(range(n))
    if n >= 2:
        import pynvml
        pynvml.nvmlInit()
        nodes = []
        for d in got:
            bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(d)).busId
            bus = bus.decode() if isinstance(bus, bytes) else bus
            try:
                nodes.append(int(open(f"/sys/bus/pci/devices/{bus[-12:].lower()}/numa_node").read()))
            except OSError:
                nodes.append(-1)
        if len(set(nodes)) == 2 and nodes.count(nodes[0]) == n // 2:
            assert all(nodes[i] != nodes[i + 1] for i in range(n - 1)), (got, nodes)
