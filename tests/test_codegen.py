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


AOT_HARNESS = r"""
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
#include "gen_aot_recon.inc"
#include "gen_aot_recon_keys.inc"
template <class G> static void run(int idx) {
  u32 x[G::K], y[G::R];
  unsigned long long s = 88172645463325252ull + idx;
  for (int it = 0; it < 512; it++) {
    for (int i = 0; i < G::K; i++) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; x[i] = (u32)(s >> 16); }
    if (it == 0) for (int i = 0; i < G::K; i++) x[i] = 0xffffffffu;
    G::combine(x, y);
    fwrite(x, 4, G::K, stdout); fwrite(y, 4, G::R, stdout);
  }
}
int main(int argc, char** argv){
  if (argc > 1) {  // dump the key table: r k coefficients...
    for (int i = 0; i < SWEC_AOT_RECON_COUNT; i++) {
      printf("%d %d", kAotReconKeys[i].r, kAotReconKeys[i].k);
      for (int j = 0; j < kAotReconKeys[i].r * kAotReconKeys[i].k; j++) printf(" %d", kAotReconKeys[i].c[j]);
      printf("\n");
    }
    return 0;
  }
#define RUN(I) run<SwecAotRecon##I>(I);
  SWEC_AOT_RECON_FOREACH(RUN)
  return 0;
}
"""


def test_compiled_in_reconstruct_matrices(tool):
    """aot_recon.cu's inputs (codegen_main.cc --aot-recon 10 4): the matrix table is exactly the fused reconstruct
    matrix of every single-shard loss and of shards 0-3 lost (oracle: rs_numpy.fused_reconstruct_rows, the statement of
    rse/src/core.rs:736-926), in that order, and every emitted combiner computes its matrix."""
    from oracle import rs_numpy as rn
    d, exe = tool
    for emit, name in (("structs", "gen_aot_recon.inc"), ("keys", "gen_aot_recon_keys.inc")):
        out = subprocess.run([exe, "--aot-recon", "10", "4", "--emit", emit], check=True, stdout=subprocess.PIPE, text=True).stdout
        (d / name).write_text(out)
    (d / "aot.cc").write_text(AOT_HARNESS)
    subprocess.run(["g++", "-O1", "-std=c++17", "-o", str(d / "aot"), str(d / "aot.cc")], check=True, cwd=d)
    keys = []
    for line in subprocess.run([str(d / "aot"), "keys"], check=True, stdout=subprocess.PIPE, text=True).stdout.splitlines():
        f = [int(v) for v in line.split()]
        keys.append(np.array(f[2:], dtype=np.uint8).reshape(f[0], f[1]))
    patterns = [(i,) for i in range(14)] + [(0, 1, 2, 3)]
    assert len(keys) == len(patterns) == 15
    for lost, key in zip(patterns, keys):
        _, outs, rows = rn.fused_reconstruct_rows(10, 4, [i not in lost for i in range(14)])
        assert list(outs) == list(lost) and (np.asarray(rows, dtype=np.uint8) == key).all(), lost
    raw = subprocess.run([str(d / "aot")], check=True, stdout=subprocess.PIPE).stdout
    words = np.frombuffer(raw, dtype="<u4")
    pos = 0
    for key in keys:
        r, k = key.shape
        blk = words[pos:pos + 512 * (k + r)].reshape(512, k + r)
        pos += 512 * (k + r)
        x = blk[:, :k].copy().view(np.uint8).reshape(-1, k, 4)
        y = blk[:, k:].copy().view(np.uint8).reshape(-1, r, 4)
        want = np.zeros_like(y)
        for p in range(r):
            for i in range(k):
                want[:, p, :] ^= rn.MUL[int(key[p, i])][x[:, i, :]]
        assert (want == y).all()
    assert pos == len(words)


XT_HARNESS = r"""
#include <cstdint>
#include <cstdio>
typedef uint32_t u32; typedef uint64_t u64;
static inline u32 ref(u32 a, u32 s){ u32 hi=a&0x80808080u; return (((a^hi)<<1) ^ ((hi>>7)*0x1du)) ^ s; }
// variant 0 of device_common.cuh, instruction by instruction: LOP, IMAD.HI, IMAD, IMAD, LOP3
static inline u32 v0(u32 a, u32 s){ u32 hi=a&0x80808080u; u32 m=(u32)(((u64)hi*0x3a000000ull)>>32); u32 a2=a*2u; u32 b2=hi*0xfffffffeu+a2; return b2^m^s; }
// variant 2: PRMT with selector 0xba98 (bytes 0-3 of a, sign-replicated), IMAD.SHL, two LOP3 0x6a = (x & c) ^ z
static inline u32 prmt_sign(u32 a){ u32 r=0; for(int b=0;b<4;b++) if((a>>(8*b))&0x80u) r|=0xffu<<(8*b); return r; }
static inline u32 v2(u32 a, u32 s){ u32 mask=prmt_sign(a); u32 a2=a*2u; u32 u=(a2&0xfefefefeu)^s; return (mask&0x1d1d1d1du)^u; }
int main(){
  u64 st=0x9E3779B97F4A7C15ull; long bad=0;
  for(long it=0; it<4000000; it++){
    st^=st<<13; st^=st>>7; st^=st<<17; u32 a=(u32)(st>>11), s=(u32)(st>>37)*2654435761u;
    if(it<65536){ a=(u32)it*0x00010001u; }            // every 16-bit pattern in both halves
    if(ref(a,s)!=v0(a,s) || ref(a,s)!=v2(a,s)) bad++;
  }
  printf("%ld\n", bad); return bad!=0;
}
"""


def test_both_multiply_by_two_spellings_equal_xtime(tmp_path):
    """The two instruction mixes of the SWAR multiply-by-2 step (device_common.cuh: variant 0 = IMAD.HI reduction
    mask + 2 IMAD shift, variant 2 = PRMT sign mask), restated instruction by instruction in C, equal the plain
    xtime on four packed bytes for 4 M random words and every 16-bit pattern."""
    (tmp_path / "xt.cc").write_text(XT_HARNESS)
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", str(tmp_path / "xt"), str(tmp_path / "xt.cc")], check=True)
    r = subprocess.run([str(tmp_path / "xt")], stdout=subprocess.PIPE, text=True)
    assert r.returncode == 0 and r.stdout.strip() == "0", r.stdout


def test_shared_power_chains_option(tool):
    """codegen.h share_powers (off by default): input-side power chains for a signal that alone occupies the high
    bit-planes of several rows.  Same bytes as the matrix product for the encode matrix, decode matrices, other ratios and
    random / degenerate matrices; fewer multiply-by-2 steps where the structure exists (encode 24 -> 20, worst-case
    decode 27 -> 21), never more."""
    from oracle import rs_numpy as rn

    def steps(gen):
        return int(gen.strip().splitlines()[-1].split("xtime_steps=")[1].split()[0])
    enc = rn.build_matrix(10, 14)[10:]
    assert steps(run_case(tool, enc)) == 24 and steps(run_case(tool, enc, extra=("--share-powers",))) == 20
    for lost in [(0, 1, 2, 3), (4, 5, 6, 7), (0, 1, 10, 11), (5,), (12,), (9, 13), (2, 6, 11, 12), (1, 5, 13)]:
        _, _, rows = rn.fused_reconstruct_rows(10, 4, [i not in lost for i in range(14)])
        a, b = steps(run_case(tool, rows)), steps(run_case(tool, rows, extra=("--share-powers",)))
        assert b <= a, (lost, a, b)
        if lost == (0, 1, 2, 3):
            assert (a, b) == (27, 21)
    rng = np.random.default_rng(4)
    for r, k in ((1, 1), (1, 10), (4, 3), (3, 17), (8, 10), (5, 32)):
        run_case(tool, rng.integers(0, 256, (r, k)), extra=("--share-powers",))
    for m in (np.zeros((2, 3)), np.eye(4), rn.build_matrix(6, 9)[6:], rn.build_matrix(20, 28)[20:],
              [[1, 2, 4, 8, 16, 32, 64, 128], [255] * 8], [[0x16] * 8 + [1, 0], [0x34] * 8 + [0, 1]]):
        run_case(tool, m, extra=("--share-powers",))
        run_case(tool, m, extra=("--share-powers", "--no-basis"))
        run_case(tool, m, extra=("--share-powers", "--no-cse"))


def test_basis_search_over_the_emitted_cost(tool):
    """`swec_codegen --search-basis`: every GF(2)-independent output basis judged by what the generator actually emits
    (steps and XORs after CSE and power sharing) instead of the a-priori estimate.  Same bytes; RS(10,4) encode goes from
    24 steps + 30 XORs (shipped) to 20 + 32 with shared power chains, shards 0,1,10,11 lost from 26 + 30 to 21 + 28."""
    from oracle import rs_numpy as rn

    def stats(gen):
        last = gen.strip().splitlines()[-1]
        return int(last.split("xtime_steps=")[1].split()[0]), int(last.split("xor_ops=")[1].split()[0])
    enc = rn.build_matrix(10, 14)[10:]
    assert stats(run_case(tool, enc)) == (24, 30)
    assert stats(run_case(tool, enc, extra=("--search-basis",))) == (24, 27)
    assert stats(run_case(tool, enc, extra=("--share-powers", "--search-basis"))) == (20, 32)
    for lost, want in (((0, 1, 10, 11), (21, 28)), ((0, 1, 2, 3), (21, 30)), ((9, 13), (14, 24))):
        _, _, rows = rn.fused_reconstruct_rows(10, 4, [i not in lost for i in range(14)])
        assert stats(run_case(tool, rows, extra=("--share-powers", "--search-basis"))) == want, lost
    rng = np.random.default_rng(5)
    for r in (2, 3, 4):
        run_case(tool, rng.integers(0, 256, (r, 10)), extra=("--share-powers", "--search-basis", "--step-cost", "4"))
