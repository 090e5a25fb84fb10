"""bench.py's reference arm (the CPU golden model on the host cores) runs without a GPU: its JSON line must carry the
keys the driver reads, and the GPU arm must refuse to run on a CPU-only box instead of falling back."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=600)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = run("--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-frames", "2", "--workload", "480p_nv12_30f_gop25_qp24")
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 0 and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["config"]["workload"] == "480p_nv12_30f_gop25_qp24"
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_gpu_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("CUDA device present")
    r = run("--steps", "1", "--warmup", "0")
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)


def test_parity_block_detects_a_single_flipped_byte():
    """bench.py's parity block: the golden model's GOP hashes against the same frames cut out of the timed stream."""
    import hashlib
    import importlib.util
    import numpy as np
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    rng = np.random.default_rng(1)
    sizes = rng.integers(50, 400, 12)
    data = rng.integers(0, 256, int(sizes.sum()), dtype=np.uint8)
    offs = np.concatenate([[0], np.cumsum(sizes)])
    gops = []
    for first in (0, 4, 8, 4):  # a repeated GOP (more cores than GOPs) is checked once
        blob = data[offs[first]:offs[first + 4]].tobytes()
        gops.append((first, 4, len(blob), hashlib.sha256(blob).hexdigest()))
    ok = bench.parity_block(gops, data, sizes)
    assert ok["equal"] and ok["gops_checked"] == 3 and ok["frames_checked"] == 12
    bad = data.copy()
    bad[int(offs[5])] ^= 1
    r = bench.parity_block(gops, bad, sizes)
    assert not r["equal"] and r["mismatching_gops_first_frame"] == [4]
