"""CPU: bench.py's stdout contract under torchrun -- exactly one JSON line, whatever C libraries (NCCL) printf to fd 1."""
import json
import os
import subprocess
import sys

from tests.util import ROOT


def test_one_json_line_on_stdout_with_noisy_fd1():
    code = (
        "import os, sys, ctypes\n"
        "sys.path.insert(0, %r)\n"
        "import bench\n"
        "bench.stdout_for_json_only(2)\n"
        "print('python-level noise')\n"
        "sys.stdout.flush()\n"
        "ctypes.CDLL(None).puts(b'NCCL version 2.28.9+cuda12.9')\n"      # what NCCL does: C stdio on fd 1
        "ctypes.CDLL(None).fflush(None)\n"
        "bench.emit({'metric': 'x', 'value': 1.5})\n" % ROOT)
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stderr
    lines = p.stdout.splitlines()
    assert len(lines) == 1 and json.loads(lines[0]) == {"metric": "x", "value": 1.5}, p.stdout
    assert "NCCL version" in p.stderr and "python-level noise" in p.stderr


def test_single_process_prints_to_stdout():
    code = "import sys\nsys.path.insert(0, %r)\nimport bench\nbench.stdout_for_json_only(1)\nbench.emit({'a': 1})\n" % ROOT
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert p.stdout.strip() == json.dumps({"a": 1})
