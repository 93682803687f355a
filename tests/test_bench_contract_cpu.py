"""bench.py's reference arm runs on host cores only, so its JSON contract can be checked here:
one JSON line on stdout, the keys the driver reads, N > 1 ranks other than 0 silent."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

HAVE_REF = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libvhp_ref_fast.so")) or \
    os.path.exists(os.path.join(ROOT, "oracle", "_build", "libvhp_oracle.so"))


def run_bench(extra_env=None, *args):
    env = dict(os.environ)
    env.update(extra_env or {})
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], cwd=ROOT, env=env,
                       capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    return p.stdout


@pytest.mark.skipif(not HAVE_REF, reason="neither oracle/_ref nor the oracle port is built")
def test_reference_arm_line():
    out = run_bench(None, "--impl", "reference", "--steps", "1", "--warmup", "0")
    lines = [l for l in out.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["unit"] == "Gcells/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["config"]["workload"].startswith("c2")
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
    assert d["gpu_launches"] == 0


@pytest.mark.skipif(not HAVE_REF, reason="neither oracle/_ref nor the oracle port is built")
def test_reference_arm_other_ranks_stay_silent():
    out = run_bench({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"},
                    "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0")
    assert out.strip() == ""
