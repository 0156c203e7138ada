"""The reference's UNCHANGED TrailNet runtime (ros/packages/caffe_ros/src/tensor_net.cpp + int8_calibrator.cpp) compiled against this
repo's NvInfer.h / NvCaffeParser.h and linked to its libraries (tools/dropin/build.sh -> dropin/_ref/caffe_ros_trailnet; ROS, Boost
and OpenCV's C++ headers are replaced by the small stand-ins of tools/dropin/include).

CPU part: the binary exists and the reference's own preprocessImage() (tensor_net.cpp:303-336), running on the cv:: stand-in, produces
the network input real OpenCV produces (tests/golden/trailnet/inputs.npz, made with cv2).  GPU part: the whole sequence of
tensor_net.cpp -- Caffe parser, buildCudaEngine, serialize, deserializeCudaEngine, execute -- on one camera frame against the
prediction the reference's test expects."""
import os
import subprocess

import numpy as np
import pytest

from tests.util import ROOT, trailnet_model_files

TN = os.path.join(ROOT, "tests", "golden", "trailnet")
EXE = os.path.join(ROOT, "dropin", "_ref", "caffe_ros_trailnet")


def _exe():
    if not os.path.exists(EXE):
        pytest.skip("dropin/_ref/caffe_ros_trailnet not built (tools/dropin/build.sh needs the reference checkout)")
    return EXE


def test_reference_preprocessing_on_the_cv_stand_in(tmp_path):
    exe = _exe()
    fr = np.load(os.path.join(TN, "frames_rgb8.npz"))
    want = np.load(os.path.join(TN, "inputs.npz"))["images"]
    for k, row in enumerate(fr["rows"]):
        raw, out = str(tmp_path / "frame.bin"), str(tmp_path / "chw.f32")
        fr["frames"][k].tofile(raw)
        h, w = fr["frames"][k].shape[:2]
        subprocess.run([exe, "preprocess", raw, str(w), str(h), out], check=True, timeout=120)
        got = np.fromfile(out, dtype=np.float32).reshape(3, 180, 320)
        err = np.abs(got - want[row])
        assert err.max() <= 5e-3 and err.mean() <= 2e-4, (row, err.max(), err.mean())      # values 0..255: fp32 summation order only


@pytest.mark.gpu
@pytest.mark.xfail(strict=False, reason="written after the round's GPU minutes were spent: the same sequence is validated through "
                                        "rt_caffe_create / rt_net_* (tests/test_gpu_trailnet.py), this binary has not run on a GPU yet")
def test_reference_tensor_net_end_to_end(tmp_path):
    exe = _exe()
    proto, model = trailnet_model_files()
    fr = np.load(os.path.join(TN, "frames_rgb8.npz"))
    exp = np.load(os.path.join(TN, "expected.npz"))["tests_cpp"]
    for k, row in enumerate(fr["rows"]):
        raw, out = str(tmp_path / "frame.bin"), str(tmp_path / "probs.f32")
        fr["frames"][k].tofile(raw)
        h, w = fr["frames"][k].shape[:2]
        r = subprocess.run([exe, "run", proto, model, raw, str(w), str(h), out], capture_output=True, text=True, timeout=120)
        print(r.stdout[-500:], r.stderr[-1500:])
        assert r.returncode == 0
        got = np.fromfile(out, dtype=np.float32)
        np.testing.assert_allclose(got, exp[row], rtol=0, atol=1e-3)          # ros/packages/caffe_ros/tests/tests.cpp:64-69,101-105
