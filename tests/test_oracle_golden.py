"""Pins the CPU oracle to the reference's own known-answer tensors.

Each case mirrors one gtest case of the reference (stereoDNN/tests/tests_main.cpp:280-1099):
same fixture files, same plugin parameters, same post-processing chain, same tolerance.
"""
import numpy as np
import pytest
import torch

from oracle import ops
from oracle.io import fixture as fx


def T(name):
    return torch.from_numpy(fx(name))


def check(actual, expected, tol):
    actual = actual.numpy() if isinstance(actual, torch.Tensor) else actual
    assert actual.shape == expected.shape, (actual.shape, expected.shape)
    if tol == "float_eq":          # EXPECT_FLOAT_EQ = within 4 ULP
        np.testing.assert_array_almost_equal_nulp(actual.astype(np.float32), expected, nulp=4)
    else:
        np.testing.assert_allclose(actual, expected, rtol=0, atol=tol)


def test_fixture_count():
    from oracle.io import fixtures
    assert len(fixtures()) == 58


# EluPluginTests.{Basic, Input4DBatchSize2}  (tests_main.cpp:280-342)
@pytest.mark.parametrize("idx", ["01", "02"])
def test_elu(idx):
    check(ops.elu(T("elu_i_" + idx)), fx("elu_o_" + idx), 1e-6)


def _conv(idx, stride, pad_start, manual_dpad=False, bias=False):
    x = T("conv3d_%s_x" % idx)
    if manual_dpad:                 # "we have to manually pad input in D dimension" (:436-439)
        x = ops.pad_d(x, 1)
    b = T("conv3d_%s_b" % idx) if bias else None
    return ops.conv3d(x, T("conv3d_%s_w" % idx), b, stride, pad_start)


# Conv3DPluginTests (tests_main.cpp:360-623); outputs pass through Transform{1,0,2,3}.
def test_conv3d_basic():
    check(ops.transform(_conv("01", (1, 1, 1), (0, 0, 0))), fx("conv3d_01_y"), 1e-5)


def test_conv3d_hw_strides():
    check(ops.transform(_conv("02", (1, 2, 2), (0, 1, 1))), fx("conv3d_02_y"), 1e-5)


def test_conv3d_dhw_strides_pad():
    check(ops.transform(_conv("03", (1, 2, 2), (0, 1, 1), manual_dpad=True)), fx("conv3d_03_y"), 1e-5)


def test_conv3d_unit_strides_sym_d():
    check(ops.transform(_conv("04", (1, 1, 1), (1, 1, 1))), fx("conv3d_04_y"), 1e-4)


def test_conv3d_asym_d():
    check(ops.transform(_conv("05", (2, 2, 2), (0, 1, 1), manual_dpad=True)), fx("conv3d_05_y"), 1e-4)


def test_conv3d_bias_elu():
    y = ops.elu(ops.transform(_conv("06", (2, 2, 2), (0, 1, 1), manual_dpad=True, bias=True)))
    check(y, fx("conv3d_06_y"), 1e-4)


def test_conv3d_multiple():
    w = T("conv3d_07_w")
    y1 = ops.transform(ops.conv3d(T("conv3d_07_x"), w, None, (1, 1, 1), (1, 1, 1)))
    y2 = ops.conv3d(ops.pad_d(y1, 1), w, None, (2, 2, 2), (0, 1, 1))
    check(ops.transform(y2), fx("conv3d_07_y"), 1e-4)


# Conv3DTransposePluginTests (tests_main.cpp:653-878).
def _xdims(name):
    return tuple(fx(name).shape[1:])


def test_conv3d_tran_basic():
    x = ops.conv3d_transpose(T("conv3d_tran_01_y"), T("conv3d_tran_01_w"), None, (1, 1, 1), (0, 0, 0),
                             _xdims("conv3d_tran_01_x"))
    # D == 1 here, so the test's trailing Transform{1,0,2,3} ([D,C,H,W]->[C,D,H,W]) keeps the flat order.
    check(x, fx("conv3d_tran_01_x"), 1e-5)


def test_conv3d_tran_hw_strides():
    x = ops.conv3d_transpose(T("conv3d_tran_02_y"), T("conv3d_tran_02_w"), None, (1, 2, 2), (0, 1, 1),
                             _xdims("conv3d_tran_02_x"))
    check(x, fx("conv3d_tran_02_x"), 1e-4)


def _tran_asym(idx, bias):
    xd = list(_xdims("conv3d_tran_%s_x" % idx))
    out_dims = [xd[0] + 1] + xd[1:]
    b = T("conv3d_tran_%s_b" % idx) if bias else None
    x = ops.conv3d_transpose(T("conv3d_tran_%s_y" % idx), T("conv3d_tran_%s_w" % idx), b,
                             (2, 2, 2), (0, 1, 1), out_dims)
    return ops.slice_d(x, 0, xd[0])


def test_conv3d_tran_asym_d():
    check(_tran_asym("03", False), fx("conv3d_tran_03_x"), 1e-4)


def test_conv3d_tran_bias_elu():
    check(ops.elu(_tran_asym("04", True)), fx("conv3d_tran_04_x"), 1e-4)


def test_conv3d_tran_multiple():
    xd = list(_xdims("conv3d_tran_05_x"))
    od1 = (9, 8, 9, 9)              # "hardcoded for now" (:838)
    od2 = [xd[0] + 1] + xd[1:]
    x1 = ops.conv3d_transpose(T("conv3d_tran_05_y"), T("conv3d_tran_05_w1"), None, (2, 2, 2), (0, 1, 1), od1)
    x1 = ops.transform(ops.slice_d(x1, 0, od1[0] - 1))
    x2 = ops.conv3d_transpose(x1, T("conv3d_tran_05_w2"), None, (2, 2, 2), (0, 1, 1), od2)
    check(ops.slice_d(x2, 0, od2[0] - 1), fx("conv3d_tran_05_x"), 1e-4)


# CostVolumePluginTests.{Basic, Large}, CorrCostVolumePluginTests.Basic (tests_main.cpp:884-990).
@pytest.mark.parametrize("idx", ["01", "02"])
def test_cost_volume(idx):
    cv = fx("cost_vol_%s_cv" % idx)
    out = ops.cost_volume(T("cost_vol_%s_l" % idx), T("cost_vol_%s_r" % idx), cv.shape[1])
    assert np.array_equal(out.numpy(), cv)          # pure copy: bit exact


def test_corr_cost_volume():
    cv = fx("corr_cost_vol_01_cv")
    out = ops.corr_cost_volume(T("corr_cost_vol_01_l"), T("corr_cost_vol_01_r"), cv.shape[1])
    check(out[:, :, None], cv, 1e-6)


# SoftargmaxPluginTests.{ArgMinBasic, ArgMinBatchSize2, ArgMaxBasic} (tests_main.cpp:1032-1099).
@pytest.mark.parametrize("idx,is_min,tol", [("01", True, 1e-6), ("02", True, 1e-5), ("03", False, 1e-6)])
def test_softargmax(idx, is_min, tol):
    check(ops.softargmax(T("softargmax_%s_x" % idx), is_min), fx("softargmax_%s_y" % idx), tol)
