"""CPU: the oracle is test infrastructure only -- nothing in the product path may import or execute it."""
import os
import re

from tests.util import ROOT


def _py_files(d):
    for dp, _, fs in os.walk(d):
        for f in fs:
            if f.endswith((".py", ".cpp", ".cu", ".cuh", ".h")):
                yield os.path.join(dp, f)


def test_product_never_touches_oracle():
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|oracle[/.]", re.M)
    for f in _py_files(os.path.join(ROOT, "redtail_b200")):
        assert not pat.search(open(f, errors="ignore").read()), f


def test_oracle_never_touches_cuda_or_product():
    for f in _py_files(os.path.join(ROOT, "oracle")):
        src = open(f).read()
        assert "redtail_b200" not in src.replace("`redtail_b200/`", ""), f
        assert ".cuda(" not in src and "device=\"cuda" not in src, f


def test_bench_uses_oracle_only_for_baselines():
    src = open(os.path.join(ROOT, "bench.py")).read()
    for m in re.finditer(r"^\s*(from|import)\s+oracle\b.*$", src, re.M):
        # every oracle import must sit inside one of the two allowed functions
        head = src[:m.start()]
        fn = re.findall(r"^def\s+(\w+)", head, re.M)[-1]
        assert fn in ("cpu_baseline", "run_reference_arm"), fn
