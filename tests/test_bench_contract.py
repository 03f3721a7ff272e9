"""CPU: the parts of bench.py's contract that need no GPU -- the reference arm (the reference's own rules on the host cores, the
one place besides the tests where oracle/ is executed), its behaviour under torchrun ranks, and the GPU arm failing loudly
without a device (there is no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, "bench.py")


def run(args, env=None):
    e = dict(os.environ)
    for k in ("WORLD_SIZE", "RANK", "LOCAL_RANK"):
        e.pop(k, None)
    e.update(env or {})
    return subprocess.run([sys.executable, BENCH] + args, capture_output=True, text=True, timeout=600, env=e)


def test_reference_arm_prints_one_contract_line():
    r = run(["--impl", "reference", "--steps", "2", "--warmup", "1"])
    assert r.returncode == 0, r.stderr[-1500:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1                                   # exactly one JSON line on stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "loa_random_playouts_per_sec" and d["unit"] == "playouts/s"
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["gpu_launches"] == 0
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and "leaves" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the same config keys the GPU arm reports (the driver compares the two dicts)
    for k in ("workload", "leaves_per_gpu", "playouts_per_leaf", "max_plies", "philox_key", "l2", "parallelism"):
        assert k in d["config"], k
    assert d["config"]["leaves_per_gpu"] == 1 << 20 and d["config"]["playouts_per_leaf"] == 16


def test_reference_arm_other_ranks_exit_quietly():
    r = run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"], env={"WORLD_SIZE": "2", "RANK": "1", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_gpu_arm_fails_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    r = run(["--steps", "1", "--warmup", "0"])
    assert r.returncode != 0 and r.stdout.strip() == ""      # no line, no silent CPU path
