"""bench.py's reference arm runs on a CPU-only machine and prints one well-formed JSON line."""

import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "clips/s" and line["value"] > 0
    assert line["higher_is_better"] is True and line["vs_baseline"] is None and line["data"] == "synthetic"
    from oracle import stage_ref

    # the unmodified reference files when __graft_entry__.build() staged them (oracle/_ref), else the oracle port
    assert line["cpu_baseline"]["kind"] == ("reference" if stage_ref.available() else "port")
    assert line["cpu_baseline"]["cores"] >= 1 and "4 clips per step" in line["cpu_baseline"]["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and "model" not in line["config"]


def test_gpus_flag_must_match_the_launch():
    """`--gpus N` is not decorative -- under a torchrun environment whose WORLD_SIZE differs it fails loudly
    (without one, bench.py re-launches itself under torch.distributed.run)."""
    env = dict(os.environ, RANK="0", WORLD_SIZE="2", LOCAL_RANK="0")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--gpus", "4", "--steps", "1"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert out.returncode != 0 and "WORLD_SIZE=2" in (out.stderr + out.stdout)


def test_reference_arm_only_rank0_prints_under_torchrun():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""
