"""The file pipeline's I/O thread pool (io_pool.h IoPool; also the bounce-copy pool of the Encoder seam) hammered under ThreadSanitizer and
AddressSanitizer on the CPU: two submitters, thousands of tiny batches — the shape
generateEcFiles("1", 50, 10000, 100) produces.  Regression test for a use-after-free of the
stack-allocated batch record (seen as a segfault on the GPU box)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

MAIN = r"""
int main() {
    IoPool pool(8, SPIN_US);
    std::atomic<long> total{0};
    std::atomic<int> bad{0};
    auto user = [&](int seed) {
        for (int it = 0; it < 6000; it++) {
            int n = 1 + (it * 7 + seed) % 14;
            std::vector<int> hit(n, 0);
            const std::function<int(int)> fn = [&](int i) { hit[i]++; total++; return (i == 3 && it % 500 == 0) ? -4 : 0; };
            int rc = pool.parallel_for(n, fn);
            for (int i = 0; i < n; i++) if (hit[i] != 1) bad++;
            if ((n > 3 && it % 500 == 0) != (rc == -4)) bad++;
        }
    };
    std::thread a(user, 1), b(user, 2);
    a.join(); b.join();
    printf("%s %ld\n", bad.load() ? "BAD" : "ok", total.load());
    return bad.load() ? 1 : 0;
}
"""


@pytest.mark.parametrize("spin_us", [0, 50])
@pytest.mark.parametrize("sanitizer", ["thread", "address"])
def test_iopool_under_sanitizers(tmp_path, sanitizer, spin_us):
    src = open(os.path.join(ROOT, "seaweedfs_b200", "csrc", "io_pool.h")).read()
    cls = src[src.index("class IoPool {"):src.index("// ---- end of IoPool")]
    head = "\n".join(f"#include <{h}>" for h in ("algorithm", "atomic", "chrono", "condition_variable", "cstdio", "deque",
                                                  "functional", "mutex", "thread", "vector"))
    (tmp_path / "t.cc").write_text(head + f"\n#define SPIN_US {spin_us}\n" + cls + MAIN)
    exe = str(tmp_path / "t")
    r = subprocess.run(["g++", "-O1", "-g", f"-fsanitize={sanitizer}", "-std=c++17", "-o", exe, str(tmp_path / "t.cc"),
                        "-lpthread"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0 and "sanitize" in r.stdout.lower() and "cannot find" in r.stdout.lower():
        pytest.skip("sanitizer runtime not installed")
    assert r.returncode == 0, r.stdout
    out = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip().startswith("ok"), out.stdout[-3000:]
