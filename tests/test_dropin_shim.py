"""CPU: the cv:: stand-in the UNCHANGED sample_app/main.cpp is compiled against (tools/dropin/include/opencv2) does what
OpenCV does for the calls main.cpp makes -- checked against cv2 itself: PNG decode, float conversion, INTER_AREA resize,
BGR->RGB, CHW, /255 (readImgFile, main.cpp:83-98) and the 16-bit PNG written from the disparity (main.cpp:317-330)."""
import os
import subprocess

import numpy as np
import pytest

from tests.util import ROOT

cv2 = pytest.importorskip("cv2")


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("shim") / "shim_check")
    subprocess.run(["g++", "-O1", "-std=c++14", "-I", os.path.join(ROOT, "tools", "dropin", "include"),
                    os.path.join(ROOT, "tools", "dropin", "shim_check.cpp"), "-o", exe, "-lz"], check=True)
    return exe


@pytest.mark.parametrize("src_hw,dst_hw,channels", [((375, 1242), (321, 1025), 3), ((97, 131), (48, 64), 3), ((60, 80), (60, 80), 4), ((50, 70), (25, 35), 1)])
def test_read_img_file_and_png16_match_opencv(shim, tmp_path, src_hw, dst_hw, channels):
    rng = np.random.default_rng(src_hw[0])
    yy, xx = np.mgrid[0:src_hw[0], 0:src_hw[1]]
    base = (127 + 90 * np.sin(xx * 0.05) * np.cos(yy * 0.03))[..., None] + rng.integers(-30, 30, src_hw + (channels,))
    img = np.clip(base, 0, 255).astype(np.uint8)
    if channels == 1:
        img = img[..., 0]
    src = str(tmp_path / "in.png")
    assert cv2.imwrite(src, img)
    h, w = dst_hw
    out_f32, out_png = str(tmp_path / "o.f32"), str(tmp_path / "o.png")
    subprocess.run([shim, src, str(w), str(h), out_f32, out_png], check=True)
    got = np.fromfile(out_f32, dtype=np.float32).reshape(3, h, w)
    ref = cv2.imread(src).astype(np.float32)                       # 8UC3 BGR like the app
    ref = cv2.resize(ref, (w, h), interpolation=cv2.INTER_AREA)
    ref = cv2.cvtColor(ref, cv2.COLOR_BGR2RGB)
    want = cv2.multiply(ref.reshape(w * h, 3).T.copy(), 1.0 / 255.0).reshape(3, h, w)
    assert np.abs(got - want).max() <= 1.2e-7
    d = (want[0] * np.float32(300.0)) * np.float32(256)
    want16 = np.clip(np.rint(d), 0, 65535).astype(np.uint16)
    got16 = cv2.imread(out_png, cv2.IMREAD_UNCHANGED)
    assert got16.dtype == np.uint16 and got16.shape == (h, w)
    assert np.abs(got16.astype(np.int32) - want16.astype(np.int32)).max() <= 1      # 1 LSB: fp32 products of 1-ulp-close inputs
