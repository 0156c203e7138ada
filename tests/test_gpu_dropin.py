"""GPU: the reference's UNCHANGED sources running on this repo's engine.

dropin/_ref/ holds binaries compiled (tools/dropin/build.sh, in the build container) from the reference's own
tests/tests_main.cpp and sample_app/*_net.cpp against this repo's headers.  Here the reference's 23-case gtest suite is
run on its own fixtures, and the reference's generated NVSmall / NVTiny builders are executed and compared with the
oracle's golden disparity.
"""
import os
import subprocess

import numpy as np
import pytest

from oracle import io as oio
from tests.util import ROOT, export_fixture_dir

pytestmark = pytest.mark.gpu

BIN = os.path.join(ROOT, "dropin", "_ref")


def _need(name):
    p = os.path.join(BIN, name)
    if not os.path.exists(p):
        pytest.skip("dropin/_ref/%s not built (tools/dropin/build.sh needs the reference checkout)" % name)
    return p


def test_reference_gtest_suite(tmp_path):
    exe = _need("nvstereo_tests")
    data = export_fixture_dir(str(tmp_path / "data"))
    # CostVolumePluginPerfTests.NVSmall moves 1 GB through the host harness; run it separately below.
    r = subprocess.run([exe, data, "--gtest_filter=*-Perf"], capture_output=True, text=True, timeout=600)
    print(r.stdout[-6000:])
    print(r.stderr[-3000:])
    assert r.returncode == 0
    assert "0 failed" in r.stdout


def test_reference_gtest_costvolume_perf_shape(tmp_path):
    exe = _need("nvstereo_tests")
    data = export_fixture_dir(str(tmp_path / "data"))
    r = subprocess.run([exe, data, "--gtest_filter=CostVolumePluginPerfTests"], capture_output=True, text=True, timeout=600)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0


# px_scale: ResNet18_2D emits sigmoid-normalised disparity; sample_app/main.cpp:325-327 multiplies it by the width.
@pytest.mark.parametrize("net,h,w,px_scale", [("nvtiny", 161, 513, 1), ("nvsmall", 321, 1025, 1),
                                              ("resnet18", 321, 1025, 1), ("resnet18_2D", 257, 513, 513)])
def test_reference_generated_builder(tmp_path, net, h, w, px_scale):
    exe = _need("nvstereo_net_driver")
    l, r = oio.load_sample_pair()
    l, r = oio.resize_pair(l, r, h, w)
    l.tofile(tmp_path / "l.bin")
    r.tofile(tmp_path / "r.bin")
    out = tmp_path / "disp.bin"
    p = subprocess.run([exe, net, str(w), str(h), oio.weights_path(net), str(tmp_path / "l.bin"), str(tmp_path / "r.bin"), str(out)],
                       capture_output=True, text=True, timeout=600)
    print(p.stdout[-2000:], p.stderr[-2000:])
    assert p.returncode == 0
    disp = np.fromfile(out, dtype=np.float32).reshape(h, w)
    gold = np.load(os.path.join(oio.GOLDEN, "disp_%s_%dx%d_f64oracle.npy" % (net, w, h)))
    err = np.abs(disp - gold) * px_scale
    print("%s: max %.3g px, mean %.3g px" % (net, err.max(), err.mean()))
    assert err.max() <= 1e-3


def test_reference_builder_plan_roundtrip(tmp_path):
    """sample_app/main.cpp:207-220,270-275: engine->serialize(), then IRuntime::deserializeCudaEngine with
    StereoDnnPluginFactory after the builder's weights, plugin container and engine are gone.  The reference can do this
    for ResNet18_2D only (its Conv3D plugins assert in serialize()); here the NVTiny plan round-trips bit-exactly."""
    exe = _need("nvstereo_net_driver")
    h, w = 161, 513
    l, r = oio.load_sample_pair()
    l, r = oio.resize_pair(l, r, h, w)
    l.tofile(tmp_path / "l.bin")
    r.tofile(tmp_path / "r.bin")
    outs = []
    for mode in ([], ["plan"]):
        out = tmp_path / ("disp%d.bin" % len(outs))
        p = subprocess.run([exe, "nvtiny", str(w), str(h), oio.weights_path("nvtiny"), str(tmp_path / "l.bin"),
                            str(tmp_path / "r.bin"), str(out)] + mode, capture_output=True, text=True, timeout=600)
        print(p.stdout[-1000:], p.stderr[-1000:])
        assert p.returncode == 0
        if mode:
            assert "engine rebuilt from it" in p.stdout
        outs.append(np.fromfile(out, dtype=np.float32))
    assert np.array_equal(outs[0], outs[1])


def _write_png8(path, chw01):
    """[3,h,w] float in [0,1] (RGB) -> 8-bit RGB PNG (stdlib zlib; the app reads it with cv::imread)."""
    import struct
    import zlib
    img = np.clip(np.rint(chw01 * 255.0), 0, 255).astype(np.uint8).transpose(1, 2, 0)
    h, w, _ = img.shape
    raw = b"".join(b"\x00" + img[y].tobytes() for y in range(h))

    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xFFFFFFFF)
    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0)) +
                chunk(b"IDAT", zlib.compress(raw, 3)) + chunk(b"IEND", b""))
    return img


@pytest.mark.parametrize("mode", ["fp32", "fp16"])
def test_reference_sample_app_unchanged(tmp_path, mode):
    """stereoDNN/sample_app/main.cpp compiled UNCHANGED (tools/dropin/build.sh -> nvstereo_sample_app): PNG images in
    (cv::imread / resize / cvtColor through the cv:: stand-in), the reference's NVTiny builder, this engine, raw disparity +
    16-bit PNG out (main.cpp:83-98,176-330).  The disparity must equal what the Python binding computes from the same
    8-bit images, and the PNG must hold round(256 * disparity)."""
    exe = _need("nvstereo_sample_app")
    import torch
    from redtail_b200 import StereoEngine
    h, w = 161, 513
    l, r = oio.load_sample_pair()
    l, r = oio.resize_pair(l, r, h, w)
    l8 = _write_png8(tmp_path / "l.png", l)
    r8 = _write_png8(tmp_path / "r.png", r)
    wpath = oio.weights_path("nvtiny")
    if mode == "fp16":
        wpath = oio.write_fp16_weights(wpath, str(tmp_path / "w16.bin"))
    out = tmp_path / "disp.bin"
    p = subprocess.run([exe, "nvsmall", str(w), str(h), wpath, str(tmp_path / "l.png"), str(tmp_path / "r.png"), str(out), mode],
                       capture_output=True, text=True, timeout=600)
    print(p.stdout[-1500:], p.stderr[-1500:])
    assert p.returncode == 0
    disp = np.fromfile(out, dtype=np.float32).reshape(h, w)
    # the same pre-processing in numpy: identity resize, RGB, CHW, * float(1/255)
    inv = np.float32(1.0 / 255.0)
    lt = torch.from_numpy(np.ascontiguousarray(l8.transpose(2, 0, 1)).astype(np.float32) * inv)[None].cuda()
    rt = torch.from_numpy(np.ascontiguousarray(r8.transpose(2, 0, 1)).astype(np.float32) * inv)[None].cuda()
    eng = StereoEngine("nvtiny", h, w, wpath, weights_dtype=mode)
    ref = eng(lt, rt).cpu().numpy()[0]
    assert np.abs(disp - ref).max() <= 1e-5
    # 16-bit PNG written by the app (cv::imwrite of the CV_16U Mat)
    try:
        import cv2
    except ImportError:
        return
    png = cv2.imread(str(out) + ".png", cv2.IMREAD_UNCHANGED)
    assert png is not None and png.dtype == np.uint16 and png.shape == (h, w)
    assert np.abs(png.astype(np.int32) - np.clip(np.rint(disp * np.float32(256)), 0, 65535).astype(np.int32)).max() <= 0
