"""bench.py's stdout contract on the CPU: the reference arm prints exactly ONE JSON line with the keys the driver reads,
even when something writes to file descriptor 1 behind Python's back (NCCL's version banner does under torchrun)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    code = ("import os, runpy, sys; sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '1', "
            "'--samples', '4096']; import bench; bench.protect_stdout(); os.write(1, b'NCCL version banner\\n'); "
            "bench.main()")
    res = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, res.stdout[:500]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "samples/s" and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "NCCL version banner" in res.stderr
