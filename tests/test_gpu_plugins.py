"""GPU parity of every plugin-level C-ABI entry point against the reference's known-answer tensors.

Each case mirrors one gtest case of the reference (stereoDNN/tests/tests_main.cpp:280-1099): same fixture, same plugin
parameters, same post-processing chain, same tolerance -- but the op under test is this repo's sm_100a kernel called
through the C-ABI (redtail_b200.ops -> libredtail_b200.so), and the expected values are the reference's own fixtures.
Additional cases compare against the fixture-pinned CPU oracle on seeded inputs at sizes the reference never tested.
"""
import numpy as np
import pytest
import torch

from oracle import ops as O
from oracle.io import fixture as fx

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def R():
    from redtail_b200 import ops
    return ops


def G(name):
    return torch.from_numpy(fx(name)).cuda()


def check(actual, expected, tol):
    a = actual.detach().float().cpu().numpy()
    e = expected.numpy() if isinstance(expected, torch.Tensor) else expected
    assert a.shape == e.shape, (a.shape, e.shape)
    if tol == "float_eq":
        np.testing.assert_array_almost_equal_nulp(a, e.astype(np.float32), nulp=4)
    else:
        np.testing.assert_allclose(a, e, rtol=0, atol=tol)


PRECS = ["simt", "fp32"]


def _prec(R, name):
    return {"simt": R.PREC_SIMT, "fp32": R.PREC_FP32, "fp16": R.PREC_FP16}[name]


def _conv(R, prec, x, w, b, stride, pad, **kw):
    # "fp32" = the tcgen05 kernel (fp16x2-split operands).  The reference's fixtures have 1..16 channels: the pack pass
    # zero-pads them to a 16-channel K block, so the tensor-core path itself meets the reference's known answers.
    op = R.Conv3d(w, b, stride, pad, tuple(x.shape[1:]), precision=_prec(R, prec), **kw)
    y = op(x)
    if prec != "simt":
        assert "umma" in R.last_kernel(), R.last_kernel()
    return y


# ---- ELU (tests_main.cpp:280-342) ----
@pytest.mark.parametrize("idx", ["01", "02"])
def test_elu(R, idx):
    check(R.elu(G("elu_i_" + idx)), fx("elu_o_" + idx), "float_eq")


def test_elu_fp16(R):
    y = R.elu(G("elu_i_01").half())
    check(y, fx("elu_o_01"), 1e-2)


def test_elu_large_unaligned(R):
    g = torch.Generator().manual_seed(1)
    x = (10 * torch.randn(3 * 1001 * 77 + 5, generator=g) - 2)
    xd = x.cuda()
    check(R.elu(xd[3:]), O.elu(x[3:]), 1e-6)      # misaligned base pointer -> scalar path


# ---- Conv3D (tests_main.cpp:360-623) ----
@pytest.mark.parametrize("prec", PRECS)
def test_conv3d_basic(R, prec):
    y = _conv(R, prec, G("conv3d_01_x"), fx("conv3d_01_w"), None, (1, 1, 1), (0, 0, 0))
    check(R.transform(y), fx("conv3d_01_y"), 1e-5)


@pytest.mark.parametrize("prec", PRECS)
def test_conv3d_hw_strides(R, prec):
    y = _conv(R, prec, G("conv3d_02_x"), fx("conv3d_02_w"), None, (1, 2, 2), (0, 1, 1))
    check(R.transform(y), fx("conv3d_02_y"), 1e-5)


@pytest.mark.parametrize("prec", PRECS)
def test_conv3d_dhw_strides_pad(R, prec):
    x = R.pad_d(G("conv3d_03_x"), 1)
    y = _conv(R, prec, x, fx("conv3d_03_w"), None, (1, 2, 2), (0, 1, 1))
    check(R.transform(y), fx("conv3d_03_y"), 1e-5)


@pytest.mark.parametrize("prec", PRECS)
def test_conv3d_unit_strides_sym_d(R, prec):
    y = _conv(R, prec, G("conv3d_04_x"), fx("conv3d_04_w"), None, (1, 1, 1), (1, 1, 1))
    check(R.transform(y), fx("conv3d_04_y"), 1e-4)


@pytest.mark.parametrize("prec", PRECS)
def test_conv3d_asym_d(R, prec):
    x = R.pad_d(G("conv3d_05_x"), 1)
    y = _conv(R, prec, x, fx("conv3d_05_w"), None, (2, 2, 2), (0, 1, 1))
    check(R.transform(y), fx("conv3d_05_y"), 1e-4)


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("fused", [False, True])
def test_conv3d_bias_elu(R, prec, fused):
    x = R.pad_d(G("conv3d_06_x"), 1)
    if fused:     # engine fusion: conv + Transform + ELU in one launch
        y = _conv(R, prec, x, fx("conv3d_06_w"), fx("conv3d_06_b"), (2, 2, 2), (0, 1, 1), fuse_elu=True, out_transposed=True)
    else:
        y = R.elu(R.transform(_conv(R, prec, x, fx("conv3d_06_w"), fx("conv3d_06_b"), (2, 2, 2), (0, 1, 1))))
    check(y, fx("conv3d_06_y"), 1e-4)


@pytest.mark.parametrize("prec", PRECS)
def test_conv3d_multiple(R, prec):
    w = fx("conv3d_07_w")
    y1 = R.transform(_conv(R, prec, G("conv3d_07_x"), w, None, (1, 1, 1), (1, 1, 1)))
    y2 = _conv(R, prec, R.pad_d(y1, 1), w, None, (2, 2, 2), (0, 1, 1))
    check(R.transform(y2), fx("conv3d_07_y"), 1e-4)


# ---- Conv3DTranspose (tests_main.cpp:653-878) ----
def _tconv(R, prec, y, w, b, stride, pad, out_dims, **kw):
    op = R.Conv3d(w, b, stride, pad, tuple(y.shape[1:]), out_dims=out_dims, transposed=True, precision=_prec(R, prec), **kw)
    if prec == "simt":
        return op

    def run(*a):
        out = op(*a)
        assert "umma" in R.last_kernel(), R.last_kernel()
        return out
    return run


@pytest.mark.parametrize("prec", PRECS)
def test_conv3d_tran_basic(R, prec):
    xd = tuple(fx("conv3d_tran_01_x").shape[1:])
    x = _tconv(R, prec, G("conv3d_tran_01_y"), fx("conv3d_tran_01_w"), None, (1, 1, 1), (0, 0, 0), xd)(G("conv3d_tran_01_y"))
    check(x, fx("conv3d_tran_01_x"), 1e-5)


@pytest.mark.parametrize("prec", PRECS)
def test_conv3d_tran_hw_strides(R, prec):
    xd = tuple(fx("conv3d_tran_02_x").shape[1:])
    x = _tconv(R, prec, G("conv3d_tran_02_y"), fx("conv3d_tran_02_w"), None, (1, 2, 2), (0, 1, 1), xd)(G("conv3d_tran_02_y"))
    check(x, fx("conv3d_tran_02_x"), 1e-4)


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("fused", [False, True])
def test_conv3d_tran_asym_d(R, prec, fused):
    xd = list(fx("conv3d_tran_03_x").shape[1:])
    od = [xd[0] + 1] + xd[1:]
    y = G("conv3d_tran_03_y")
    if fused:
        x = _tconv(R, prec, y, fx("conv3d_tran_03_w"), None, (2, 2, 2), (0, 1, 1), od, slice_d=1)(y)
    else:
        x = R.slice_d(_tconv(R, prec, y, fx("conv3d_tran_03_w"), None, (2, 2, 2), (0, 1, 1), od)(y), 0, xd[0])
    check(x, fx("conv3d_tran_03_x"), 1e-4)


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("fused", [False, True])
def test_conv3d_tran_bias_elu(R, prec, fused):
    xd = list(fx("conv3d_tran_04_x").shape[1:])
    od = [xd[0] + 1] + xd[1:]
    y = G("conv3d_tran_04_y")
    if fused:
        x = _tconv(R, prec, y, fx("conv3d_tran_04_w"), fx("conv3d_tran_04_b"), (2, 2, 2), (0, 1, 1), od, slice_d=1, fuse_elu=True)(y)
    else:
        x = R.elu(R.slice_d(_tconv(R, prec, y, fx("conv3d_tran_04_w"), fx("conv3d_tran_04_b"), (2, 2, 2), (0, 1, 1), od)(y), 0, xd[0]))
    check(x, fx("conv3d_tran_04_x"), 1e-4)


@pytest.mark.parametrize("prec", PRECS)
def test_conv3d_tran_multiple(R, prec):
    xd = list(fx("conv3d_tran_05_x").shape[1:])
    od1 = (9, 8, 9, 9)
    od2 = [xd[0] + 1] + xd[1:]
    y = G("conv3d_tran_05_y")
    x1 = _tconv(R, prec, y, fx("conv3d_tran_05_w1"), None, (2, 2, 2), (0, 1, 1), od1, slice_d=1)(y)
    x1 = R.transform(x1)
    x2 = _tconv(R, prec, x1, fx("conv3d_tran_05_w2"), None, (2, 2, 2), (0, 1, 1), od2, slice_d=1)(x1)
    check(x2, fx("conv3d_tran_05_x"), 1e-4)


def test_conv3d_tran_skip_fusion(R):
    """deconv -> Slice -> +skip -> ELU in one launch == the four separate ops (oracle)."""
    g = torch.Generator().manual_seed(3)
    y = torch.randn(2, 16, 3, 5, 7, generator=g)
    w = torch.randn(16, 3, 8, 3, 3, generator=g) * 0.2
    b = torch.randn(8, generator=g)
    od = (7, 8, 9, 13)
    skip = torch.randn(2, 6, 8, 9, 13, generator=g)
    ref = O.elu(O.slice_d(O.conv3d_transpose(y, w, b, (2, 2, 2), (0, 1, 1), od), 0, 6) + skip)
    op = R.Conv3d(w.numpy(), b.numpy(), (2, 2, 2), (0, 1, 1), (16, 3, 5, 7), out_dims=od, transposed=True,
                  precision=R.PREC_SIMT, slice_d=1, fuse_elu=True)
    check(op(y.cuda(), skip.cuda()), ref, 1e-4)


# ---- cost volume (tests_main.cpp:884-1026) ----
@pytest.mark.parametrize("idx", ["01", "02"])
def test_cost_volume(R, idx):
    cv = fx("cost_vol_%s_cv" % idx)
    out = R.cost_volume(G("cost_vol_%s_l" % idx), G("cost_vol_%s_r" % idx), cv.shape[1])
    assert R.last_kernel() == "cost_volume_tma"
    assert np.array_equal(out.cpu().numpy(), cv)      # pure data movement: bit exact


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
@pytest.mark.parametrize("shape,disp", [((2, 5, 37, 131), 17), ((1, 3, 8, 4099), 9), ((1, 2, 3, 5), 7), ((1, 4, 161, 513), 48)])
def test_cost_volume_shapes(R, dtype, shape, disp):
    g = torch.Generator().manual_seed(7)
    l, r = torch.randn(shape, generator=g).to(dtype), torch.randn(shape, generator=g).to(dtype)
    out = R.cost_volume(l.cuda(), r.cuda(), disp)
    assert torch.equal(out.cpu(), O.cost_volume(l, r, disp))


def test_cost_volume_nvsmall_full_size(R):
    """CostVolumePluginPerfTests.NVSmall shape (tests_main.cpp:938-958), with value checks via size-independent
    properties: left half is D identical copies; right half plane d equals plane 0 shifted by d with zero fill."""
    g = torch.Generator().manual_seed(11)
    l = torch.randn(1, 32, 161, 513, generator=g).cuda()
    r = torch.randn(1, 32, 161, 513, generator=g).cuda()
    cv = R.cost_volume(l, r, 48)
    assert cv.shape == (1, 48, 64, 161, 513)
    assert torch.equal(cv[:, :, :32], l[:, None].expand(-1, 48, -1, -1, -1))
    for d in (0, 1, 7, 47):
        assert torch.equal(cv[:, d, 32:, :, d:], r[:, :, :, : 513 - d])
        assert not cv[:, d, 32:, :, :d].any()


def test_corr_cost_volume(R):
    cv = fx("corr_cost_vol_01_cv")
    out = R.corr_cost_volume(G("corr_cost_vol_01_l"), G("corr_cost_vol_01_r"), cv.shape[1])
    check(out[:, :, None], cv, 1e-6)


def test_corr_cost_volume_fp16(R):
    cv = fx("corr_cost_vol_01_cv")
    out = R.corr_cost_volume(G("corr_cost_vol_01_l").half(), G("corr_cost_vol_01_r").half(), cv.shape[1])
    check(out[:, :, None], cv, 1e-2)


def test_corr_cost_volume_large(R):
    g = torch.Generator().manual_seed(5)
    l, r = torch.randn(2, 32, 33, 257, generator=g), torch.randn(2, 32, 33, 257, generator=g)
    check(R.corr_cost_volume(l.cuda(), r.cuda(), 48), O.corr_cost_volume(l, r, 48), 1e-4)


# ---- softargmax (tests_main.cpp:1032-1099) ----
@pytest.mark.parametrize("idx,is_min,tol", [("01", True, 2e-6), ("02", True, 1e-5), ("03", False, 2e-6)])
def test_softargmax(R, idx, is_min, tol):
    check(R.softargmax(G("softargmax_%s_x" % idx), is_min), fx("softargmax_%s_y" % idx), tol)


def test_softargmax_nvsmall_size(R):
    g = torch.Generator().manual_seed(2)
    x = torch.randn(1, 96, 1, 321, 1025, generator=g) * 3
    check(R.softargmax(x.cuda(), True), O.softargmax(x, True), 2e-4)


# ---- data movement + TRT-native layers ----
def test_pad_slice_transform_concat(R):
    g = torch.Generator().manual_seed(4)
    x = torch.randn(2, 5, 3, 7, 9, generator=g)
    xd = x.cuda()
    assert torch.equal(R.pad_d(xd, 2).cpu(), O.pad_d(x, 2))
    assert torch.equal(R.slice_d(xd, 1, 4).cpu(), O.slice_d(x, 1, 4))
    assert torch.equal(R.transform(xd).cpu(), O.transform(x))
    a, b = torch.randn(2, 3, 11, 13, generator=g), torch.randn(2, 1, 11, 13, generator=g)
    assert torch.equal(R.concat_channels(a.cuda(), b.cuda()).cpu(), torch.cat([a, b], 1))
    check(R.eltwise_sum(xd, xd), x + x, 0)
    check(R.sigmoid(xd), torch.sigmoid(x), 1e-6)
    check(R.scale(xd, 0.5, 2.0, 1.0), x * 2 + 0.5, 1e-6)
    check(R.convert(R.convert(xd, torch.float16), torch.float32), x.half().float(), 0)


@pytest.mark.parametrize("k,stride,pad,cin,cout", [(5, 2, 2, 3, 32), (3, 1, 1, 32, 32), (3, 1, 1, 32, 8), (3, 2, 1, 33, 64)])
def test_conv2d(R, k, stride, pad, cin, cout):
    g = torch.Generator().manual_seed(6)
    x = torch.randn(2, cin, 37, 65, generator=g)
    w = torch.randn(cout, cin, k, k, generator=g) * 0.1
    b = torch.randn(cout, generator=g)
    op = R.Conv2d(w.numpy(), b.numpy(), (stride, stride), (pad, pad), (37, 65), fuse_elu=True)
    check(op(x.cuda()), O.elu(O.conv2d(x, w, b, (stride, stride), (pad, pad))), 2e-4)


def test_deconv2d(R):
    g = torch.Generator().manual_seed(8)
    x = torch.randn(2, 16, 17, 33, generator=g)
    w = torch.randn(16, 8, 3, 3, generator=g) * 0.1
    b = torch.randn(8, generator=g)
    op = R.Conv2d(w.numpy(), b.numpy(), (2, 2), (1, 1), (17, 33), transposed=True)
    check(op(x.cuda()), O.deconv2d(x, w, b, (2, 2), (1, 1)), 2e-4)


# ---- NVSmall-class conv shapes vs oracle (sizes the reference never value-checked) ----
@pytest.mark.parametrize("prec,tol", [("simt", 2e-4), ("fp32", 2e-4), ("fp16", 5e-2)])
@pytest.mark.parametrize("cin,cout,stride", [(64, 32, 1), (32, 64, 2), (64, 64, 1), (128, 128, 1)])
def test_conv3d_nvsmall_class(R, prec, tol, cin, cout, stride):
    g = torch.Generator().manual_seed(cin + cout)
    d, h, w_ = (7, 19, 37) if stride == 1 else (8, 19, 37)
    x = torch.randn(1, d + (stride == 2), cin, h, w_, generator=g)
    if stride == 2:
        x[:, -1] = 0                                   # the PaddingPlugin plane
    w = torch.randn(cout, 3, cin, 3, 3, generator=g) * (1.0 / np.sqrt(27 * cin))
    b = torch.randn(cout, generator=g)
    pad = (1, 1, 1) if stride == 1 else (0, 1, 1)
    ref = O.elu(O.transform(O.conv3d(x.double(), w.double(), b.double(), (stride,) * 3, pad))).float()
    y = _conv(R, prec, x.cuda(), w.numpy(), b.numpy(), (stride,) * 3, pad, fuse_elu=True, out_transposed=True)
    check(y, ref, tol)
    if prec != "simt":
        assert "tc" in R.last_kernel() or "umma" in R.last_kernel(), R.last_kernel()


@pytest.mark.parametrize("cin,cout,stride,wdtype", [(64, 64, 1, np.float16), (32, 32, 1, np.float32), (32, 64, 2, np.float16), (128, 128, 1, np.float32)])
def test_conv3d_fp16_exact_weights(R, cin, cout, stride, wdtype):
    """The reference's fp16 configuration (trt_weights_fp16.bin): every weight is an fp16 value, so the A_hi x W_lo product of
    the fp32-split scheme is identically zero and the kernel issues two products instead of three -- same tolerance as
    the fp32 path, whether the weights arrive as an fp16 array or as fp32 values that are fp16-representable."""
    g = torch.Generator().manual_seed(cin + 7 * cout)
    d, h, w_ = (5, 19, 37) if stride == 1 else (6, 19, 37)
    x = torch.randn(1, d + (stride == 2), cin, h, w_, generator=g)
    if stride == 2:
        x[:, -1] = 0
    w = (torch.randn(cout, 3, cin, 3, 3, generator=g) * (1.0 / np.sqrt(27 * cin))).half()
    b = torch.randn(cout, generator=g).half()
    pad = (1, 1, 1) if stride == 1 else (0, 1, 1)
    ref = O.elu(O.transform(O.conv3d(x.double(), w.double(), b.double(), (stride,) * 3, pad))).float()
    y = _conv(R, "fp32", x.cuda(), w.numpy().astype(wdtype), b.numpy().astype(wdtype), (stride,) * 3, pad, fuse_elu=True, out_transposed=True)
    assert R.last_kernel() == "conv3d_umma_fp16x2split_w16", R.last_kernel()
    check(y, ref, 2e-4)


@pytest.mark.parametrize("prec,tol", [("simt", 2e-4), ("fp32", 2e-4), ("fp16", 5e-2)])
@pytest.mark.parametrize("cin,cout", [(128, 64), (64, 32), (32, 1)])
def test_conv3d_transpose_nvsmall_class(R, prec, tol, cin, cout):
    g = torch.Generator().manual_seed(cin * 3 + cout)
    y = torch.randn(1, cin, 4, 9, 17, generator=g)
    w = torch.randn(cin, 3, cout, 3, 3, generator=g) * (1.0 / np.sqrt(27 * cin / 8))
    b = torch.randn(cout, generator=g)
    od = (9, cout, 17, 33)
    ref = O.slice_d(O.conv3d_transpose(y.double(), w.double(), b.double(), (2, 2, 2), (0, 1, 1), od), 0, 8).float()
    op = _tconv(R, prec, y.cuda(), w.numpy(), b.numpy(), (2, 2, 2), (0, 1, 1), od, slice_d=1)
    check(op(y.cuda()), ref, tol)


# ---- engine-internal RT_LAYOUT_SPLIT16 (channels-last fp16 hi/lo): same values as the dense plugin layouts ----
def test_split16_roundtrip_and_cost_volume(R):
    g = torch.Generator().manual_seed(21)
    x = torch.randn(2, 5, 16, 9, 37, generator=g) * 30
    s16 = R.dense_to_split16(x.cuda())
    back = R.split16_to_dense(s16)
    assert (back.cpu() - x).abs().max() <= 30 * 4 * 2 ** -22          # 22 significant bits
    l, r = torch.randn(2, 8, 11, 70, generator=g), torch.randn(2, 8, 11, 70, generator=g)
    cv = R.split16_to_dense(R.cost_volume_split16(l.cuda(), r.cuda(), 13))
    check(cv, O.cost_volume(l, r, 13), 4 * 2 ** -20)


@pytest.mark.parametrize("cin,cout,stride", [(64, 32, 1), (32, 64, 2), (128, 128, 1)])
def test_conv3d_split16_chain(R, cin, cout, stride):
    """conv in split16 -> split16 (incl. the virtual Padding plane for the stride-2 layers) == dense plugin chain."""
    g = torch.Generator().manual_seed(cin + 3 * cout)
    d, h, w_ = (6, 19, 37) if stride == 1 else (8, 19, 37)
    x = torch.randn(1, d, cin, h, w_, generator=g)
    w = torch.randn(cout, 3, cin, 3, 3, generator=g) * (1.0 / np.sqrt(27 * cin))
    b = torch.randn(cout, generator=g)
    pad = (1, 1, 1) if stride == 1 else (0, 1, 1)
    xin = O.pad_d(x, 1) if stride == 2 else x
    ref = O.elu(O.transform(O.conv3d(xin.double(), w.double(), b.double(), (stride,) * 3, pad))).float()
    op = R.Conv3d(w.numpy(), b.numpy(), (stride,) * 3, pad, tuple(x.shape[1:]), precision=R.PREC_FP32, fuse_elu=True,
                  in_layout=R.LAYOUT_SPLIT16, out_layout=R.LAYOUT_SPLIT16, pad_end_d=1 if stride == 2 else 0)
    y = R.split16_to_dense(op(R.dense_to_split16(x.cuda())))
    check(y, ref, 2e-4)


@pytest.mark.parametrize("cout,d,h,w_,n,wdtype", [(32, 6, 19, 37, 1, np.float32), (32, 3, 8, 16, 2, np.float32), (24, 4, 9, 50, 1, np.float32),
                                                   (32, 5, 21, 33, 1, np.float16), (32, 2, 30, 17, 1, np.float32)])
def test_conv3d_depth_stationary(R, monkeypatch, cout, d, h, w_, n, wdtype):
    """32 -> <=32 channel 3x3x3 stride-1 convs in split16 run on the depth-stationary tcgen05 kernel (conv3d_ds.cu): against the
    oracle, and bit for bit against the generic kernel forced to the same accumulation chunks (one chunk per filter plane and
    column) -- both kernels add the same products in the same order."""
    g = torch.Generator().manual_seed(cout + d + 5 * h)
    x = torch.randn(n, d, 32, h, w_, generator=g)
    w = (torch.randn(cout, 3, 32, 3, 3, generator=g) * (1.0 / np.sqrt(27 * 32)))
    b = torch.randn(cout, generator=g)
    if wdtype == np.float16:
        w, b = w.half().float(), b.half().float()
    ref = O.elu(O.transform(O.conv3d(x.double(), w.double(), b.double(), (1, 1, 1), (1, 1, 1)))).float()
    xs = R.dense_to_split16(x.cuda())
    mk = lambda: R.Conv3d(w.numpy(), b.numpy(), (1, 1, 1), (1, 1, 1), tuple(x.shape[1:]), precision=R.PREC_FP32, fuse_elu=True,
                          in_layout=R.LAYOUT_SPLIT16, out_layout=R.LAYOUT_SPLIT16)
    ys = mk()(xs)
    assert R.last_kernel() == ("conv3d_ds_fp16x2split_w16" if wdtype == np.float16 else "conv3d_ds_fp16x2split"), R.last_kernel()
    check(R.split16_to_dense(ys), ref, 2e-4)
    monkeypatch.setenv("REDTAIL_TC_DS", "0")
    monkeypatch.setenv("REDTAIL_TC_CHAIN", "6")
    yg = mk()(xs)
    assert "umma" in R.last_kernel(), R.last_kernel()
    assert torch.equal(ys, yg), float((R.split16_to_dense(ys) - R.split16_to_dense(yg)).abs().max())


def test_conv3d_transpose_split16_skip(R):
    g = torch.Generator().manual_seed(77)
    y = torch.randn(1, 64, 4, 9, 17, generator=g)
    w = torch.randn(64, 3, 32, 3, 3, generator=g) * 0.05
    b = torch.randn(32, generator=g)
    od = (9, 32, 17, 33)
    skip = torch.randn(1, 8, 32, 17, 33, generator=g)
    ref = O.elu(O.slice_d(O.conv3d_transpose(y.double(), w.double(), b.double(), (2, 2, 2), (0, 1, 1), od), 0, 8) + skip.double()).float()
    op = R.Conv3d(w.numpy(), b.numpy(), (2, 2, 2), (0, 1, 1), (64, 4, 9, 17), out_dims=od, transposed=True, precision=R.PREC_FP32,
                  slice_d=1, fuse_elu=True, in_layout=R.LAYOUT_SPLIT16, out_layout=R.LAYOUT_SPLIT16)
    yin = R.dense_to_split16(y.permute(0, 2, 1, 3, 4).contiguous().cuda())     # [N,D,K,H,W] -> split16 [D][H][W][K]
    out = R.split16_to_dense(op(yin, R.dense_to_split16(skip.cuda())))
    check(out, ref, 3e-4)


@pytest.mark.parametrize("d,h,w_,n,slice_d,is_min,wdtype", [(6, 9, 17, 1, 1, True, np.float32), (4, 21, 40, 2, 1, True, np.float32),
                                                            (5, 8, 16, 1, 0, False, np.float32), (7, 13, 33, 1, 1, True, np.float16)])
def test_deconv_softargmax_fused(R, d, h, w_, n, slice_d, is_min, wdtype):
    """Conv3DTranspose (32 -> 1, stride 2) + Slice + Softargmax as one kernel (deconv_softargmax.cu) against the oracle's three ops."""
    g = torch.Generator().manual_seed(d * 100 + h)
    y = torch.randn(n, 32, d, h, w_, generator=g)
    w = torch.randn(32, 3, 1, 3, 3, generator=g) * 0.25
    b = torch.randn(1, generator=g)
    if wdtype == np.float16:
        w, b = w.half().float(), b.half().float()
    od = (2 * d + 1, 1, 2 * h - 1, 2 * w_ - 1)
    vol = O.slice_d(O.conv3d_transpose(y.double(), w.double(), b.double(), (2, 2, 2), (0, 1, 1), od), 0, od[0] - slice_d)
    ref = O.softargmax(vol[:, :, 0], is_min).float()
    op = R.Conv3d(w.numpy(), b.numpy(), (2, 2, 2), (0, 1, 1), (32, d, h, w_), out_dims=od, transposed=True, precision=R.PREC_FP32,
                  slice_d=slice_d, in_layout=R.LAYOUT_SPLIT16, fuse_softargmax=1 if is_min else 2)
    yin = R.dense_to_split16(y.permute(0, 2, 1, 3, 4).contiguous().cuda())
    out = op(yin)
    assert R.last_kernel().startswith("deconv_softargmax"), R.last_kernel()
    check(out, ref.reshape(out.shape), 2e-4)


# ---- fused CostVolume -> Conv3D (the volume is never built): same values as the unfused plugin pair ----
@pytest.mark.parametrize("prec,tol", [("simt", 2e-4), ("fp32", 2e-4)])
@pytest.mark.parametrize("n,c,k,h,w_,disp", [(1, 32, 32, 11, 70, 13), (2, 16, 16, 5, 9, 5), (1, 32, 8, 7, 131, 48)])
def test_costvol_conv3d_fused(R, prec, tol, n, c, k, h, w_, disp):
    g = torch.Generator().manual_seed(c + k + w_)
    l, r = torch.randn(n, c, h, w_, generator=g), torch.randn(n, c, h, w_, generator=g)
    w = torch.randn(k, 3, 2 * c, 3, 3, generator=g) * (1.0 / np.sqrt(27 * 2 * c))
    b = torch.randn(k, generator=g)
    cv = O.cost_volume(l, r, disp)                                              # [N,D,2C,H,W]
    ref = O.elu(O.transform(O.conv3d(cv.double(), w.double(), b.double(), (1, 1, 1), (1, 1, 1)))).float()   # [N,D,K,H,W]
    p = {"simt": R.PREC_SIMT, "fp32": R.PREC_FP32}[prec]
    op = R.CostVolumeConv3d(w.numpy(), b.numpy(), (c, h, w_), disp, precision=p, fuse_elu=True, out_transposed=True)
    check(op(l.cuda(), r.cuda()), ref, tol)
    assert R.last_kernel() == "cvconv_combine_kernel"
    op = R.CostVolumeConv3d(w.numpy(), b.numpy(), (c, h, w_), disp, precision=p, fuse_elu=False, out_transposed=False)
    ref_nt = O.conv3d(cv.double(), w.double(), b.double(), (1, 1, 1), (1, 1, 1)).float()                      # [N,K,D,H,W]
    check(op(l.cuda(), r.cuda()), ref_nt, tol)
    op = R.CostVolumeConv3d(w.numpy(), b.numpy(), (c, h, w_), disp, precision=p, fuse_elu=True, out_layout=R.LAYOUT_SPLIT16)
    check(R.split16_to_dense(op(l.cuda(), r.cuda())), ref, tol + 4 * 2 ** -20)
