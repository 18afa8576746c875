"""CPU: bench.py's contract pieces that do not need a GPU -- stdout carries nothing but the JSON line (libraries that
print to stdout are diverted), and without a GPU the product arm fails loudly instead of printing a number."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_stdout_is_reserved_for_the_json_line():
    code = (
        "import os, sys; sys.path.insert(0, %r); import bench\n"
        "bench.claim_stdout()\n"
        "os.write(1, b'NCCL version banner\\n')\n"      # what a library writing to fd 1 does
        "print('python print to stdout')\n"
        "bench.emit({'metric': 'Msamples/s', 'value': 1.0})\n" % ROOT
    )
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, check=True)
    lines = out.stdout.splitlines()
    assert len(lines) == 1 and json.loads(lines[0]) == {"metric": "Msamples/s", "value": 1.0}
    assert "NCCL version banner" in out.stderr and "python print to stdout" in out.stderr


def test_product_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0"], capture_output=True, text=True)
    assert out.returncode != 0
    assert out.stdout.strip() == ""


def test_wave_count_helper():
    sys.path.insert(0, ROOT)
    import bench
    assert bench.JOB_WAVES(2) == 1 and bench.JOB_WAVES(32) == 16
