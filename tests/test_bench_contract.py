"""bench.py's reference arm on a tiny bounded sample: the JSON line carries every key the driver reads (it is the one arm
that runs without a GPU, so the contract can be checked here), and the arm really computes the benchmarked sweep: the
unmodified reference, driven on reference model objects that hold the synthetic workload's cluster states, returns the
scores, SNR statistics, labels and counts of the oracle's sweep over the same sample."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _have_reference():
    from oracle import refshim
    return refshim.find_reference() is not None


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-beats", "1024", "--cpu-ref-beats", "512"], capture_output=True, text=True, cwd=ROOT, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "beats/s" and d["higher_is_better"] is True and d["dtype"] == "f64"
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["vs_baseline"] is None and d["data"] == "synthetic"
    cb = d["cpu_baseline"]
    n = 512 if cb["kind"] == "reference" else 1024
    assert d["value"] > 0 and abs(d["value"] - n / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]
    assert "workload" in d["config"] and "model" not in d["config"]
    assert cb["kind"] == ("reference" if _have_reference() else "port")
    assert cb["cores"] >= 1 and cb["value"] == d["value"] and f"{n} beats" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "beats/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                         capture_output=True, text=True, cwd=ROOT, env=env, timeout=300)
    assert out.returncode == 0 and not [l for l in out.stdout.splitlines() if l.startswith("{")]


@pytest.mark.skipif(not _have_reference(), reason="reference package neither mounted nor staged")
def test_reference_arm_computes_the_benchmarked_sweep():
    import bench
    ref, port = {}, {}
    bench.reference_sweep_beats_per_s(192, 64, 2, 6, steps=1, outputs=ref)
    bench.cpu_sweep_beats_per_s(192, 64, 2, 6, steps=1, outputs=port)
    assert np.max(np.abs(ref["q"] - port["q"]) / np.abs(port["q"])) < 1e-10
    assert np.max(np.abs(ref["snr"] - port["snr"])) < 1e-9
    assert np.array_equal(ref["z"], port["z"])
    assert np.array_equal(ref["transStateCount"], port["transStateCount"])
