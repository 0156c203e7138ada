"""Fixture / weight / input readers for the oracle and the tests (TEST INFRASTRUCTURE).

Formats follow the reference:
  * plugin fixtures: int32 ndims | int32 dims[] | float32 data
    (writer scripts/test_data_generator.py:34-39, reader tests/tests_main.cpp:259-275),
    re-packed by tests/golden/make_golden.py into plugin_fixtures.npz;
  * weight files: cstring name, u32 count, count x (f32|f16)
    (writer scripts/tensorrt_model_builder.py:52-60, reader sample_app/main.cpp:111-134).
"""
import os
import struct

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

_fixtures = None


def fixtures():
    global _fixtures
    if _fixtures is None:
        with np.load(os.path.join(GOLDEN, "plugin_fixtures.npz")) as z:
            _fixtures = {k: z[k] for k in z.files}
    return _fixtures


def fixture(name):
    return fixtures()[name]


def read_weights(path, dtype=np.float32):
    """-> dict name -> flat numpy array (shapes are not stored; they come from the builder)."""
    out = {}
    esz = np.dtype(dtype).itemsize
    with open(path, "rb") as f:
        raw = f.read()
    i = 0
    while i < len(raw):
        j = raw.index(b"\0", i)
        name = raw[i:j].decode()
        (cnt,) = struct.unpack_from("<I", raw, j + 1)
        out[name] = np.frombuffer(raw, dtype=dtype, count=cnt, offset=j + 5).copy()
        i = j + 5 + cnt * esz
    return out


def write_fp16_weights(src_fp32, dst):
    """The reference's fp16 weight file of a net: same entries, payloads rounded to fp16 (the reference's generator writes
    both files from the same arrays, scripts/tensorrt_model_builder.py:52-60; byte identity with
    models/*/TensorRT/trt_weights_fp16.bin is checked by tests/golden/make_golden_fp16.py and pinned by fp16_md5.json)."""
    w = read_weights(src_fp32)
    with open(dst, "wb") as f:
        for name, a in w.items():
            f.write(name.encode() + b"\0")
            f.write(struct.pack("<I", a.size))
            f.write(a.astype("<f2").tobytes())
    return dst


def weights_path(net, dtype="fp32"):
    return os.path.join(GOLDEN, "weights", "%s_%s.bin" % (net, dtype))


def load_sample_pair():
    """The reference's sample stereo pair (sample_app/data/img_{left,right}.bin), [3,321,1025] f32 in [0,1]."""
    l = np.load(os.path.join(GOLDEN, "images", "kitti_left_1025x321.f16.npy")).astype(np.float32)
    r = np.load(os.path.join(GOLDEN, "images", "kitti_right_1025x321.f16.npy")).astype(np.float32)
    return l, r


def resize_pair(left, right, h, w):
    """Area-resample a [3,H,W] pair to (h, w) (the apps use INTER_AREA, sample_app/main.cpp:90)."""
    import torch
    import torch.nn.functional as F
    if left.shape[1:] == (h, w):
        return left, right
    out = []
    for img in (left, right):
        t = torch.from_numpy(img)[None]
        out.append(F.adaptive_avg_pool2d(t, (h, w))[0].numpy().astype(np.float32))
    return out[0], out[1]


def synthetic_pair(h, w, seed=1234):
    """Seeded KITTI-shaped synthetic pair (set S2): smooth random texture, right image
    = left warped by a piecewise-planar disparity in [2, 80*w/1025] px, so the cost volume has structure."""
    rng = np.random.default_rng(seed)
    yy, xx = np.meshgrid(np.arange(h, dtype=np.float32), np.arange(w, dtype=np.float32), indexing="ij")
    img = np.zeros((3, h, w), np.float32)
    for c in range(3):
        acc = np.zeros((h, w), np.float32)
        for _ in range(6):
            fx, fy = rng.uniform(0.01, 0.35, 2)
            ph = rng.uniform(0, 2 * np.pi)
            acc += rng.uniform(0.3, 1.0) * np.sin(fx * xx + fy * yy + ph)
        acc = (acc - acc.min()) / (acc.max() - acc.min() + 1e-6)
        img[c] = acc
    left = np.clip(img + 0.05 * rng.uniform(0, 1, img.shape).astype(np.float32), 0, 1).astype(np.float32)
    dmax = 80.0 * w / 1025.0
    disp = np.where(yy > h * 0.55, 2.0 + (dmax - 2.0) * (yy - h * 0.55) / (h * 0.45), 2.0 + 0.25 * dmax * xx / w)
    xs = np.clip(xx + disp, 0, w - 1)           # right[x] = left[x + d]
    x0 = np.floor(xs).astype(np.int64)
    x1 = np.minimum(x0 + 1, w - 1)
    a = (xs - x0).astype(np.float32)
    rows = np.arange(h)[:, None]
    right = (1 - a) * left[:, rows, x0] + a * left[:, rows, x1]
    return left, np.ascontiguousarray(right, dtype=np.float32)
