"""The straight-line GF(2^8) generator (csrc/codegen.cc) checked on the CPU: the emitted combine()
is compiled as plain C++ (macros mapped to scalar code) and compared byte-for-byte with the oracle's
multiplication table for the RS(10,4) encode matrix, decode matrices and random matrices."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "seaweedfs_b200", "csrc")

HARNESS = r"""
#include <cstdint>
#include <cstdio>
typedef uint32_t u32;
#define __device__
#define __forceinline__ inline
#define SWEC_X2(a,b) ((a)^(b))
#define SWEC_X3(a,b,c) ((a)^(b)^(c))
static inline u32 xt(u32 a){ u32 hi=a&0x80808080u; return ((a^hi)<<1) ^ ((hi>>7)*0x1du); }
#define SWEC_XT0A(a) xt(a)
#define SWEC_XT0B(a) xt(a)
#define SWEC_XT1A(a,s) (xt(a)^(s))
#define SWEC_XT1B(a,s) (xt(a)^(s))
#include "gen.inc"
int main(){
  u32 x[G::K], y[G::R];
  unsigned long long s = 88172645463325252ull;
  for (int it = 0; it < 4096; it++) {
    for (int i = 0; i < G::K; i++) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; x[i] = (u32)(s >> 16); }
    if (it == 0) for (int i = 0; i < G::K; i++) x[i] = 0xffffffffu;
    if (it == 1) for (int i = 0; i < G::K; i++) x[i] = 0x80018001u << (i & 3);
    G::combine(x, y);
    fwrite(x, 4, G::K, stdout); fwrite(y, 4, G::R, stdout);
  }
  return 0;
}
"""


@pytest.fixture(scope="module")
def tool(tmp_path_factory):
    d = tmp_path_factory.mktemp("codegen")
    exe = str(d / "swec_codegen")
    subprocess.run(["g++", "-O1", "-std=c++17", "-o", exe] + [os.path.join(CSRC, f) for f in
                   ("codegen_main.cc", "codegen.cc", "gf256.cc")], check=True)
    (d / "harness.cc").write_text(HARNESS)
    return d, exe


def run_case(tool, rows, extra=()):
    from oracle import rs_numpy as rn
    d, exe = tool
    rows = np.asarray(rows, dtype=np.uint8)
    r, k = rows.shape
    gen = subprocess.run([exe, "--name", "G", *extra, "--rows", str(r), str(k)] + [str(int(v)) for v in rows.ravel()],
                         check=True, stdout=subprocess.PIPE, text=True).stdout
    (d / "gen.inc").write_text(gen)
    subprocess.run(["g++", "-O1", "-std=c++17", "-o", str(d / "h"), str(d / "harness.cc")], check=True, cwd=d)
    raw = subprocess.run([str(d / "h")], check=True, stdout=subprocess.PIPE).stdout
    words = np.frombuffer(raw, dtype="<u4").reshape(-1, k + r)
    x = words[:, :k].copy().view(np.uint8).reshape(-1, k, 4)
    y = words[:, k:].copy().view(np.uint8).reshape(-1, r, 4)
    want = np.zeros_like(y)
    for p in range(r):
        for i in range(k):
            want[:, p, :] ^= rn.MUL[int(rows[p, i])][x[:, i, :]]
    assert (want == y).all()
    return gen


def test_rs10_4_encode_matrix(tool):
    from oracle import rs_numpy as rn
    gen = run_case(tool, rn.build_matrix(10, 14)[10:])
    stats = gen.strip().splitlines()[-1]
    steps = int(stats.split("xtime_steps=")[1].split()[0])
    xors = int(stats.split("xor_ops=")[1].split()[0])
    assert steps <= 28 and xors <= 40, stats      # vs 28 steps + 76 XORs unoptimised
    run_case(tool, rn.build_matrix(10, 14)[10:], extra=("--no-basis",))
    run_case(tool, rn.build_matrix(10, 14)[10:], extra=("--no-basis", "--no-cse"))


@pytest.mark.parametrize("erased", [(0, 1, 2, 3), (0, 1, 10, 11), (5,), (9, 13), (2, 6, 11, 12)])
def test_decode_matrices(tool, erased):
    from oracle import rs_numpy as rn
    _, _, rows = rn.fused_reconstruct_rows(10, 4, [i not in erased for i in range(14)])
    run_case(tool, rows)


def test_random_and_degenerate_matrices(tool):
    rng = np.random.default_rng(2)
    for r, k in ((1, 1), (1, 10), (4, 3), (3, 17), (8, 10), (5, 32)):
        run_case(tool, rng.integers(0, 256, (r, k)))
    run_case(tool, np.zeros((2, 3)))
    run_case(tool, np.eye(4))
    run_case(tool, [[1, 2, 4, 8, 16, 32, 64, 128], [255] * 8])
