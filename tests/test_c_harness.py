"""The C ABI called from plain C, the way cgo's C side calls it: tests/c/cgo_shaped_harness.c
is compiled with gcc -std=c11 -pedantic -Werror against include/swec.h (so the header is checked as C, not
C++), linked with libswec.so only, and run — host-only calls here, the Encode/Reconstruct/ReconstructData
call shapes of ec_encoder.go:265,360 and store_ec.go:551 on the GPU box, parity compared with the oracle."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c", "cgo_shaped_harness.c")


@pytest.fixture(scope="module")
def harness(swec, tmp_path_factory):
    from seaweedfs_b200 import _native
    libdir = os.path.dirname(_native.library_path())
    exe = str(tmp_path_factory.mktemp("charness") / "cgo_shaped_harness")
    cmd = ["gcc", "-std=c11", "-O1", "-Wall", "-Wextra", "-Werror", "-pedantic", "-D_POSIX_C_SOURCE=200809L",
           "-I", os.path.join(ROOT, "include"), SRC, "-o", exe, "-L", libdir, "-l:libswec.so",
           "-Wl,-rpath," + libdir]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    return exe


def test_c_caller_host_only_calls(harness):
    r = subprocess.run([harness, "abi"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120)
    assert r.returncode == 0, r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("n", [256 * 1024, 1 << 20, 4099])
def test_c_caller_encode_reconstruct(cuda, harness, oracle, tmp_path, n):
    """n = 256 KiB is the reference's production batch (ec_encoder.go:68); 4099 exercises the byte tail."""
    rng = np.random.default_rng(n)
    data = rng.integers(0, 256, 10 * n, dtype=np.uint8)
    data.tofile(tmp_path / "in.bin")
    r = subprocess.run([harness, "encode", str(tmp_path / "in.bin"), str(tmp_path / "out.bin"), str(n)],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0, r.stdout
    got = np.fromfile(tmp_path / "out.bin", dtype=np.uint8).reshape(4, n)
    want = oracle.encode(10, 4, [data[i * n:(i + 1) * n] for i in range(10)])
    for p in range(4):
        assert (got[p] == want[p]).all(), f"parity {p} from the C caller differs from the oracle"
