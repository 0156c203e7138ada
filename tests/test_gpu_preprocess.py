"""GPU: the image side of the apps against OpenCV's own results.

rt_preprocess_bgr8 replaces readImgFile (sample_app/main.cpp:83-98); the known answers in tests/golden/preprocess_cv2.npz
were produced by cv2 with exactly that call sequence (tests/golden/make_golden_preprocess.py).  Tolerance: 2 ulp of the
[0,1] output -- the kernel forms the same fp32 sums in the same order as cv::resize."""
import os

import numpy as np
import pytest
import torch

from oracle import io as oio

pytestmark = pytest.mark.gpu


def test_preprocess_matches_opencv_known_answers():
    from redtail_b200 import ops
    z = np.load(os.path.join(oio.GOLDEN, "preprocess_cv2.npz"))
    n = len([k for k in z.files if k.startswith("src_")])
    assert n >= 5
    for i in range(n):
        src, want = z["src_%d" % i], z["dst_%d" % i]
        got = ops.preprocess_bgr8(torch.from_numpy(src[None]).cuda(), want.shape[1], want.shape[2]).cpu().numpy()[0]
        assert ops.last_kernel() == "preprocess_bgr8_area"
        err = np.abs(got - want).max()
        print(i, src.shape, want.shape, "max err %.3g" % err)
        assert err <= 2.4e-7, (i, err)


def test_preprocess_kitti_size_batch_and_live_opencv():
    """1242x375 -> 1025x321 (the apps' case), batch of 2, compared with cv2 run on the spot when it is importable."""
    cv2 = pytest.importorskip("cv2")
    from redtail_b200 import ops
    rng = np.random.default_rng(9)
    yy, xx = np.mgrid[0:375, 0:1242]
    imgs = []
    for k in range(2):
        base = 127 + 100 * np.sin(xx * (0.01 + 0.003 * k))[..., None] * np.cos(yy * 0.02)[..., None] * np.ones(3)
        imgs.append(np.clip(base + rng.integers(-20, 20, (375, 1242, 3)), 0, 255).astype(np.uint8))
    src = np.stack(imgs)
    got = ops.preprocess_bgr8(torch.from_numpy(src).cuda(), 321, 1025).cpu().numpy()
    for k in range(2):
        img = cv2.resize(src[k].astype(np.float32), (1025, 321), interpolation=cv2.INTER_AREA)
        img = cv2.cvtColor(img, cv2.COLOR_BGR2RGB)
        want = cv2.multiply(img.reshape(1025 * 321, 3).T.copy(), 1.0 / 255.0).reshape(3, 321, 1025)
        assert np.abs(got[k] - want).max() <= 2.4e-7


def test_disparity_to_u16_and_png(tmp_path):
    from redtail_b200 import ops
    d = torch.tensor([[0.0, 0.5, 1.001953125, 94.99, 255.998, 300.0, -3.0]], device="cuda")
    u = ops.disparity_to_u16(d, 256.0).cpu().numpy()
    assert u.tolist() == [[0, 128, 256, 24317, 65535, 65535, 0]]          # cvRound + saturate_cast<ushort>
    ops.write_png16(tmp_path / "d.png", u)
    cv2 = pytest.importorskip("cv2")
    assert np.array_equal(cv2.imread(str(tmp_path / "d.png"), cv2.IMREAD_UNCHANGED), u)


def test_preprocess_rejects_upscaling():
    from redtail_b200 import ops
    with pytest.raises(ops.RedtailError):
        ops.preprocess_bgr8(torch.zeros((1, 8, 8, 3), dtype=torch.uint8, device="cuda"), 16, 16)


def test_engine_execute_images_end_to_end(tmp_path):
    """rt_stereo_execute_images: 8-bit BGR host images -> GPU pre-processing -> NVTiny -> disparity + 16-bit PNG payload, equal
    to running the same steps one by one (ops.preprocess_bgr8, the engine, ops.disparity_to_u16)."""
    from redtail_b200 import StereoEngine, ops
    h, w = 161, 513
    l, r = oio.load_sample_pair()                                   # [3,321,1025] RGB in [0,1]
    to_bgr8 = lambda a: np.ascontiguousarray(np.clip(np.rint(a[::-1].transpose(1, 2, 0) * 255.0), 0, 255).astype(np.uint8))
    lb, rb = torch.from_numpy(to_bgr8(l)[None]), torch.from_numpy(to_bgr8(r)[None])    # 321x1025 sources, resized 2:1 (area)
    eng = StereoEngine("nvtiny", h, w, oio.weights_path("nvtiny"))
    disp = torch.empty((1, h, w), dtype=torch.float32).pin_memory()
    u16 = torch.empty((1, h, w), dtype=torch.uint16).pin_memory()
    eng.execute_images(lb.pin_memory(), rb.pin_memory(), out=disp, out_u16=u16)
    lt, rt = ops.preprocess_bgr8(lb.cuda(), h, w), ops.preprocess_bgr8(rb.cuda(), h, w)
    ref = eng(lt, rt)
    assert np.abs(disp.numpy() - ref.cpu().numpy()).max() <= 1e-6
    assert np.array_equal(u16.numpy(), ops.disparity_to_u16(ref).cpu().numpy())
    assert 1.0 < float(ref.mean()) < 60.0                           # a real disparity map, not zeros
    ops.write_png16(tmp_path / "disp.png", u16.numpy()[0])
    assert os.path.getsize(tmp_path / "disp.png") > 2 * h * w
