"""`bench.py --impl reference` — the reference's own CPU mixer timed on the host cores — runs without
a GPU; this checks its JSON line against the driver's contract (same metric/unit/config as the CUDA
arm, impl/cpu_baseline/e2e keys) so the round-end ratio can be computed."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.ref
def test_reference_arm_prints_one_contract_line():
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                                   "--steps", "2", "--warmup", "1"], text=True, cwd=ROOT, timeout=600)
    lines = [l for l in out.splitlines() if l.strip().startswith("{")]
    assert len(lines) == 1, out
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    assert d["metric"].startswith("voice-samples/s") and d["unit"] == "voice-samples/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"]
    assert e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"]


@pytest.mark.ref
def test_reference_arm_only_rank0_works_under_torchrun_env():
    """Under torchrun (N > 1) rank 0 alone runs and prints the line; the other ranks exit 0 silently."""
    env = dict(os.environ)
    env.update(RANK="1", LOCAL_RANK="1", WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29999")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--warmup", "0"], text=True, cwd=ROOT, timeout=600, env=env,
                       capture_output=True)
    assert p.returncode == 0, p.stderr[-500:]
    assert not [l for l in p.stdout.splitlines() if l.strip().startswith("{")], p.stdout
