"""The oracle's tensor-library ops against plain Python loops written from the reference's index formulas (tiny cases).

CPU only.  The fixtures of the reference pin the oracle on the shapes the reference tested; these loops pin the index
arithmetic itself (padding, stride, transposed-conv gather, disparity shift) independently of torch's convolution code.
Formulas: lib/kernels.cu:50-97 (cost volume), lib/conv_utils.cpp:14-81 + conv3d_plugin.cpp:187-216 (conv3d, TF cross-correlation,
symmetric pad), conv3d_transpose_plugin.cpp:205-243 (transposed conv = gradient of that conv), softargmax_plugin.cpp:167-205."""
import numpy as np
import torch

from oracle import ops


def _rng(seed):
    return np.random.default_rng(seed)


def test_cost_volume_loops():
    g = _rng(1)
    c, h, w, d_max = 3, 4, 7, 5
    l, r = g.standard_normal((1, c, h, w)), g.standard_normal((1, c, h, w))
    ref = np.zeros((1, d_max, 2 * c, h, w))
    for d in range(d_max):
        for ch in range(c):
            for y in range(h):
                for x in range(w):
                    ref[0, d, ch, y, x] = l[0, ch, y, x]
                    ref[0, d, c + ch, y, x] = r[0, ch, y, x - d] if x >= d else 0.0
    out = ops.cost_volume(torch.from_numpy(l), torch.from_numpy(r), d_max).numpy()
    assert np.array_equal(out, ref)
    corr = np.zeros((1, d_max, h, w))
    for d in range(d_max):
        for y in range(h):
            for x in range(d, w):
                corr[0, d, y, x] = sum(l[0, ch, y, x] * r[0, ch, y, x - d] for ch in range(c))
    out = ops.corr_cost_volume(torch.from_numpy(l), torch.from_numpy(r), d_max).numpy()
    np.testing.assert_allclose(out, corr, rtol=0, atol=1e-12)


def _conv3d_loops(x, w, b, stride, pad):
    n, d, c, h, wd = x.shape
    k, v, _, r, s = w.shape
    do = (d + 2 * pad[0] - v) // stride[0] + 1
    ho = (h + 2 * pad[1] - r) // stride[1] + 1
    wo = (wd + 2 * pad[2] - s) // stride[2] + 1
    y = np.zeros((n, k, do, ho, wo))
    for kk in range(k):
        for od in range(do):
            for oh in range(ho):
                for ow in range(wo):
                    acc = b[kk]
                    for vv in range(v):
                        for cc in range(c):
                            for rr in range(r):
                                for ss in range(s):
                                    id_, ih, iw = od * stride[0] - pad[0] + vv, oh * stride[1] - pad[1] + rr, ow * stride[2] - pad[2] + ss
                                    if 0 <= id_ < d and 0 <= ih < h and 0 <= iw < wd:
                                        acc += w[kk, vv, cc, rr, ss] * x[0, id_, cc, ih, iw]
                    y[0, kk, od, oh, ow] = acc
    return y


def test_conv3d_loops():
    g = _rng(2)
    for stride, pad, d in (((1, 1, 1), (1, 1, 1), 4), ((2, 2, 2), (0, 1, 1), 5), ((2, 2, 2), (1, 1, 1), 4)):
        x = g.standard_normal((1, d, 2, 5, 6))
        w = g.standard_normal((3, 3, 2, 3, 3))
        b = g.standard_normal(3)
        out = ops.conv3d(torch.from_numpy(x), torch.from_numpy(w), torch.from_numpy(b), stride, pad).numpy()
        np.testing.assert_allclose(out, _conv3d_loops(x, w, b, stride, pad), rtol=0, atol=1e-12)


def test_conv3d_transpose_loops():
    """x[dx,c,hx,wx] = b[c] + sum over (k,v,r,s,dy,hy,wy) with dx + pd = dy*sd + v (and likewise H, W) of w[k,v,c,r,s] * y[k,dy,hy,wy]."""
    g = _rng(3)
    for pad, od in (((0, 1, 1), (7, 2, 9, 11)), ((1, 1, 1), (6, 2, 9, 11))):
        y = g.standard_normal((1, 3, 3, 5, 6))
        w = g.standard_normal((3, 3, 2, 3, 3))
        b = g.standard_normal(2)
        st = (2, 2, 2)
        ref = np.zeros((1, od[0], od[1], od[2], od[3]))
        for c in range(od[1]):
            ref[0, :, c] = b[c]
        for k in range(3):
            for dy in range(3):
                for hy in range(5):
                    for wy in range(6):
                        for v in range(3):
                            for r in range(3):
                                for s in range(3):
                                    dx, hx, wx = dy * st[0] + v - pad[0], hy * st[1] + r - pad[1], wy * st[2] + s - pad[2]
                                    if 0 <= dx < od[0] and 0 <= hx < od[2] and 0 <= wx < od[3]:
                                        ref[0, dx, :, hx, wx] += w[k, v, :, r, s] * y[0, k, dy, hy, wy]
        out = ops.conv3d_transpose(torch.from_numpy(y), torch.from_numpy(w), torch.from_numpy(b), st, pad, od).numpy()
        np.testing.assert_allclose(out, ref, rtol=0, atol=1e-12)


def test_softargmax_and_elu_loops():
    g = _rng(4)
    x = g.standard_normal((1, 6, 3, 4)) * 3
    for is_min in (False, True):
        ref = np.zeros((1, 1, 3, 4))
        for yy in range(3):
            for xx in range(4):
                z = -x[0, :, yy, xx] if is_min else x[0, :, yy, xx]
                e = np.exp(z - z.max())
                ref[0, 0, yy, xx] = float((e / e.sum() * np.arange(6)).sum())
        np.testing.assert_allclose(ops.softargmax(torch.from_numpy(x), is_min).numpy(), ref, rtol=0, atol=1e-12)
    v = g.standard_normal(50)
    np.testing.assert_allclose(ops.elu(torch.from_numpy(v)).numpy(), np.where(v > 0, v, np.expm1(v)), rtol=0, atol=1e-15)
