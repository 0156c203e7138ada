"""GPU: the TrailNet S-ResNet-18 classifier (BASELINE config C5) through the nvcaffeparser1-compatible
parser and the engine, against the reference's own expected predictions (ros/packages/caffe_ros/tests/tests.cpp:64-69, 1e-3)
and the float64 Caffe oracle."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from tests.util import trailnet_model_files

TN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "trailnet")
PROTO, MODEL = trailnet_model_files()


@pytest.fixture(scope="module")
def data():
    return np.load(os.path.join(TN, "inputs.npz"))["images"], np.load(os.path.join(TN, "expected.npz"))


def test_trailnet_reference_predictions(data):
    from redtail_b200 import CaffeNet
    x, exp = data
    net = CaffeNet(PROTO, MODEL, "out", max_batch=5)
    assert net.in_chw == (3, 180, 320) and net.out_chw == (6, 1, 1)
    y = net(torch.from_numpy(x).cuda()).cpu().numpy().reshape(5, 6)
    np.testing.assert_allclose(y, exp["tests_cpp"], rtol=0, atol=1e-3)          # the reference's test, its tolerance
    np.testing.assert_allclose(y, exp["oracle_f64"], rtol=0, atol=2e-4)         # the float64 oracle
    # one image at a time (the reference node runs batch 1) gives the same rows
    y1 = np.stack([net(torch.from_numpy(x[i:i + 1]).cuda()).cpu().numpy().reshape(6) for i in range(5)])
    np.testing.assert_allclose(y1, y, rtol=0, atol=1e-6)


def test_trailnet_plan_roundtrip_and_host_path(data):
    """tensor_net.cpp:168-217: build, serialize, deserialize (the model cache file), execute with host buffers."""
    from redtail_b200 import CaffeNet
    x, exp = data
    net = CaffeNet(PROTO, MODEL, "out", max_batch=2)
    plan = net.serialize()
    y0 = net(torch.from_numpy(x[:2]).cuda()).cpu().numpy()
    net.close()
    net2 = CaffeNet.deserialize(plan, max_batch=8)
    xb = np.concatenate([x, x[:3]])
    h_in, h_out = torch.from_numpy(xb).pin_memory(), torch.empty((8, 6, 1, 1)).pin_memory()
    net2.execute_host(h_in, h_out)
    y = h_out.numpy().reshape(8, 6)
    np.testing.assert_array_equal(y[:2], y0.reshape(2, 6))
    np.testing.assert_allclose(y[:5], exp["tests_cpp"], rtol=0, atol=1e-3)
    np.testing.assert_array_equal(y[5:], y[:3])
    with pytest.raises(Exception):
        CaffeNet.deserialize(plan[:1000])


def test_classifier_layers_vs_oracle():
    """Caffe pooling (ceil-mode extent, clipped windows, padded AVE divisor), S-ReLU chain, InnerProduct, Softmax kernels."""
    import ctypes as C
    from redtail_b200._lib import kernels_lib
    from oracle import caffe
    lib = kernels_lib()
    g = torch.Generator().manual_seed(5)
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    P = lambda t: C.c_void_p(t.data_ptr())
    x = torch.randn(3, 5, 87, 157, generator=g)
    for kind, k, st, pd in (("MAX", 3, 2, 0), ("AVE", 3, 1, 0), ("AVE", 3, 2, 1), ("MAX", 2, 2, 1)):
        ref = caffe._pool(x.double().numpy(), kind, k, st, pd)
        y = torch.empty(ref.shape, dtype=torch.float32, device="cuda")
        rc = lib.rt_pool2d(P(x.cuda()), P(y), 3, 5, 87, 157, ref.shape[2], ref.shape[3], k, st, pd, int(kind == "MAX"), s)
        assert rc == 0
        np.testing.assert_allclose(y.cpu().numpy(), ref, rtol=0, atol=2e-6)
    c = 5
    s1, b1, s2, b2 = [torch.randn(c, generator=g) for _ in range(4)]
    xd = x.cuda()
    y = torch.empty_like(xd)
    assert lib.rt_srelu(P(xd), P(y), 3, c, 87 * 157, P(s1.cuda()), P(b1.cuda()), P(s2.cuda()), P(b2.cuda()), s) == 0
    v = lambda t: t.view(1, -1, 1, 1)
    ref = torch.relu(x * v(s1) + v(b1)) * v(s2) + v(b2)
    np.testing.assert_allclose(y.cpu().numpy(), ref.numpy(), rtol=0, atol=2e-6)
    assert lib.rt_scale_channel(P(xd), P(y), 3, c, 87 * 157, P(s1.cuda()), None, s) == 0
    np.testing.assert_allclose(y.cpu().numpy(), (x * v(s1)).numpy(), rtol=0, atol=1e-6)
    w, b = torch.randn(3, 4096, generator=g) * 0.02, torch.randn(3, generator=g)
    xi = torch.randn(7, 4096, generator=g)
    yo = torch.empty(7, 3, device="cuda")
    assert lib.rt_fully_connected(P(xi.cuda()), P(w.cuda()), P(b.cuda()), P(yo), 7, 4096, 3, s) == 0
    np.testing.assert_allclose(yo.cpu().numpy(), (xi.double() @ w.double().t() + b.double()).numpy(), rtol=0, atol=2e-5)
    sm = torch.empty_like(yo)
    assert lib.rt_softmax_channels(P(yo), P(sm), 7, 3, 1, s) == 0
    np.testing.assert_allclose(sm.cpu().numpy(), torch.softmax(yo.cpu().double(), dim=1).numpy(), rtol=0, atol=1e-6)


@pytest.mark.parametrize("cin,cout,stride,h,w_,n", [(64, 64, 1, 43, 78, 2), (128, 256, 1, 22, 39, 3), (256, 256, 2, 22, 39, 2), (64, 128, 2, 11, 20, 1)])
def test_conv2d_more_than_128_outputs_and_fused_srelu(cin, cout, stride, h, w_, n):
    """2-D convolutions of the TrailNet blocks on the tcgen05 kernel: more than 128 output channels as <= 128-channel parts, residual add
    and the S-ReLU chain (Scale -> ReLU -> Scale, per channel) in the epilogue -- against the float64 oracle."""
    from redtail_b200 import ops as R
    from oracle import ops as O
    g = torch.Generator().manual_seed(cin + cout + stride)
    x = torch.randn(n, 1, cin, h, w_, generator=g)
    wt = torch.randn(cout, 1, cin, 3, 3, generator=g) * (1.0 / np.sqrt(9 * cin))
    b = torch.randn(cout, generator=g)
    act = torch.randn(4, cout, generator=g)
    ref = O.conv3d(x.double(), wt.double(), b.double(), (1, stride, stride), (0, 1, 1))          # [N, K, 1, Ho, Wo]
    skip = torch.randn(ref.shape, generator=g).float()
    v = lambda t: t.double().view(1, -1, 1, 1, 1)
    ref_act = torch.relu((ref + skip.double()) * v(act[0]) + v(act[1])) * v(act[2]) + v(act[3])
    op = R.Conv3d(wt.numpy(), b.numpy(), (1, stride, stride), (0, 1, 1), (1, cin, h, w_), precision=R.PREC_FP32, act=act.numpy())
    y = op(x.cuda(), skip=skip.cuda())
    assert "umma" in R.last_kernel(), R.last_kernel()
    np.testing.assert_allclose(y.cpu().numpy(), ref_act.float().numpy(), rtol=0, atol=3e-4)
    op2 = R.Conv3d(wt.numpy(), b.numpy(), (1, stride, stride), (0, 1, 1), (1, cin, h, w_), precision=R.PREC_FP32)
    np.testing.assert_allclose(op2(x.cuda()).cpu().numpy(), ref.float().numpy(), rtol=0, atol=3e-4)


def test_im2col_conv_7x7():
    """TrailNet conv1 as im2col + 1x1 tensor-core convolution == the plain 7x7 stride-2 convolution."""
    import ctypes as C
    from redtail_b200 import ops as R
    from redtail_b200._lib import kernels_lib
    from oracle import ops as O
    g = torch.Generator().manual_seed(9)
    n, c, h, w_, k = 2, 3, 45, 80, 64
    x = torch.rand(n, c, h, w_, generator=g) * 255
    wt = torch.randn(k, c, 7, 7, generator=g) * 0.01
    b = torch.randn(k, generator=g)
    ref = O.conv2d(x.double(), wt.double(), b.double(), (2, 2), (0, 0))
    ho, wo, kp = ref.shape[2], ref.shape[3], 192
    m = torch.empty((n, 2, 1, ho, wo, kp), dtype=torch.float16, device="cuda")
    rc = kernels_lib().rt_im2col_split16(C.c_void_p(x.cuda().data_ptr()), C.c_void_p(m.data_ptr()), n, c, h, w_, 7, 7, 2, 0, ho, wo, kp,
                                          C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0
    w2 = torch.zeros(k, 1, kp, 1, 1)
    w2[:, 0, :147, 0, 0] = wt.reshape(k, 147)
    op = R.Conv3d(w2.numpy(), b.numpy(), (1, 1, 1), (0, 0, 0), (1, kp, ho, wo), precision=R.PREC_FP32, in_layout=R.LAYOUT_SPLIT16)
    y = op(m)                                                   # [N, K, 1, Ho, Wo]
    np.testing.assert_allclose(y.cpu().numpy()[:, :, 0], ref.float().numpy(), rtol=0, atol=2e-3)      # values up to ~30
