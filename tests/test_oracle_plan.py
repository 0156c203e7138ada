"""The graph-level oracle (oracle/plan.py): plan parser + interpreter over the fixture-pinned ops.

CPU only.  The NVTiny plan is dumped host-side from the reference's unchanged generated builder (tools/dropin net driver,
`dump` mode -- no GPU involved) and must reproduce the hand-written oracle (oracle/nets.py, stored golden) to rounding;
that pins the interpreter that produces the ResNet-18 / ResNet18_2D goldens (tests/golden/make_golden.py)."""
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle import io as oio, ops, plan as P

DRIVER = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "dropin", "_ref", "nvstereo_net_driver")


def test_slab_convolutions_equal_plain_ops(monkeypatch):
    monkeypatch.setattr(P, "_SLAB_BYTES", 2e4)            # force several slabs on these small tensors
    g = torch.Generator().manual_seed(0)
    for stride, pad, d in (((1, 1, 1), (1, 1, 1), 7), ((2, 2, 2), (0, 1, 1), 9), ((2, 2, 2), (1, 1, 1), 8)):
        x = torch.randn(1, d, 4, 9, 11, generator=g, dtype=torch.float64)
        w = torch.randn(5, 3, 4, 3, 3, generator=g, dtype=torch.float64)
        b = torch.randn(5, generator=g, dtype=torch.float64)
        assert torch.equal(ops.conv3d(x, w, b, stride, pad), P._conv3d_slabs(x, w, b, stride, pad))
    for pad, od in (((0, 1, 1), (9, 3, 17, 21)), ((1, 1, 1), (8, 3, 17, 21)), ((0, 1, 1), (8, 3, 17, 21))):
        y = torch.randn(1, 6, 4, 9, 11, generator=g, dtype=torch.float64)
        w = torch.randn(6, 3, 3, 3, 3, generator=g, dtype=torch.float64)
        b = torch.randn(3, generator=g, dtype=torch.float64)
        ref = ops.conv3d_transpose(y, w, b, (2, 2, 2), pad, od)
        assert (ref - P._conv3d_transpose_slabs(y, w, b, (2, 2, 2), pad, od)).abs().max() < 1e-13


@pytest.mark.skipif(not os.path.exists(DRIVER), reason="tools/dropin not built (needs the reference sources)")
def test_nvtiny_plan_reproduces_handwritten_oracle(tmp_path):
    h, w = 161, 513
    z = tmp_path / "z.bin"
    np.zeros(3 * h * w, dtype=np.float32).tofile(z)
    out = tmp_path / "nvtiny.plan"
    p = subprocess.run([DRIVER, "nvtiny", str(w), str(h), oio.weights_path("nvtiny"), str(z), str(z), str(out), "dump"],
                       capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    plan = P.parse(out.read_bytes())
    assert len(plan["layers"]) == 59 and [n for _, n, _ in plan["inputs"]] == ["left", "right"]
    l, r = oio.load_sample_pair()
    l, r = oio.resize_pair(l, r, h, w)
    res = P.execute(plan, {"left": l[None].astype(np.float64), "right": r[None].astype(np.float64)})
    disp = res["disp"].reshape(h, w)
    gold = np.load(os.path.join(oio.GOLDEN, "disp_nvtiny_513x161_f64oracle.npy"))
    assert np.abs(disp - gold).max() < 5e-6
