"""bench.py's contract with the driver, as far as it can be checked without a GPU: the reference arm runs on the
host cores alone, prints ONE JSON line with the keys the tier contract names, and non-zero ranks of a torchrun
launch stay silent; our own arm refuses to run without a device instead of falling back."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, stdout=subprocess.PIPE,
                          stderr=subprocess.PIPE, text=True, timeout=600, env=e, cwd=ROOT)


def test_reference_arm_json_line():
    r = _run(["--impl", "reference", "--gpus", "1", "--steps", "2", "--warmup", "1"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "rs10_4_encode_input_GBps" and d["unit"] == "GB/s"
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1 and d["higher_is_better"] is True
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["dtype"] == "u8" and d["data"] == "synthetic"
    assert d["vs_baseline"] is None and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_only_rank_zero_speaks():
    r = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"], env={"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_own_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    r = _run(["--steps", "1", "--warmup", "3"])
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)


def test_batch_golden_is_self_consistent_and_reproducible_on_a_sample(oracle=None):
    """tests/golden/batch256.json (the CPU oracle's digests of BASELINE configs[3]'s 256 seeded 30 GiB volumes, written
    by tests/golden/make_batch_golden.py) is what bench.py's `batch` leg must reproduce on the GPU at every N.  Here:
    the committed per-volume digests combine to the committed batch digest (placement-independent rule of
    seaweedfs_b200/sharding.py), bench.py finds it, and a miniature batch recomputed now with the same script
    logic gives the same per-volume numbers as a direct oracle call."""
    sys.path.insert(0, ROOT)
    import bench
    from oracle import pyoracle as po
    from seaweedfs_b200 import sharding
    with open(os.path.join(ROOT, "tests", "golden", "batch256.json")) as f:
        g = json.load(f)
    per = {v: int(x, 16) for v, x in enumerate(g["per_volume"])}
    assert len(per) == g["volumes"] == 256
    assert "%016x" % sharding.combine_digests(per) == g["digest"] == g["prefix_digests"]["256"]
    for n in (1, 2, 8, 64):
        assert "%016x" % sharding.combine_digests({v: per[v] for v in range(n)}) == g["prefix_digests"][str(n)]
    assert bench.load_batch_golden(256, 30 << 30)[0] == g["digest"]
    assert bench.load_batch_golden(256, 1 << 30)[0] is None                      # other volume size: no golden
    # the fold rule of bench.batch_leg / make_batch_golden.py on a small volume, against digests of the encoded image
    small = 25 * (1 << 20) + 13
    d14 = po.volume_digests(small, sharding.volume_seed(3))
    want = [po.np_digest(s) for s in po.encode_dat_image(po.synth(0, small, sharding.volume_seed(3)))]
    assert d14 == want
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_batch_golden
    fold = 0
    for one in want[10:]:
        fold = (fold * 0x100000001B3 + one) & ((1 << 64) - 1)
    assert make_batch_golden.fold_parity(d14) == fold
    assert int(g["shard_digests_volume0"][10], 16) and len(g["shard_digests_volume0"]) == 14
