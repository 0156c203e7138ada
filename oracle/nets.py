"""Network-level CPU oracle: the reference's NVSmall / NVTiny stereo nets (TEST INFRASTRUCTURE).

Wiring follows the reference's generated builders layer by layer
(`sample_app/nvsmall_1025x321_net.cpp:21-427`, `sample_app/nvtiny_513x161_net.cpp`) using the
fixture-pinned ops of `oracle/ops.py`, with the reference's trained weights.
"""
import numpy as np
import torch

from . import ops

# name -> (tower widths conv1..conv5, max_disp at half res, 3-D encoder widths conv3D_1..8, decoder widths)
SPECS = {
    # nvsmall_1025x321_net.cpp:48-165 (towers), :168-171 (cost_vol D=48), :174-323, :331-415
    "nvsmall": dict(tower=(32, 32, 32, 32, 32), max_disp=48,
                    enc=(32, 32, 64, 64, 64, 128, 128, 128), dec=(64, 32, 1)),
    # nvtiny_513x161_net.cpp:48-160, :168, :175-319, :331-402
    "nvtiny": dict(tower=(32, 32, 32, 32, 8), max_disp=24,
                   enc=(16, 16, 32, 32, 32, 64, 64, 64), dec=(32, 16, 1)),
}


def _t(a, dtype):
    return torch.from_numpy(np.asarray(a)).to(dtype)


def tower(wts, side, x, spec, dtype):
    """conv1 5x5 s2 p2 + ELU, conv2..4 3x3 + ELU, conv5 3x3 (no activation)."""
    x = ops.scale(x, float(wts[side + "_scale_shift"][0]), float(wts[side + "_scale_scale"][0]),
                  float(wts[side + "_scale_power"][0]))
    cin = x.shape[1]
    for i, cout in enumerate(spec["tower"], start=1):
        k = 5 if i == 1 else 3
        w = _t(wts["%s_conv%d_k" % (side, i)], dtype).view(cout, cin, k, k)
        b = _t(wts["%s_conv%d_b" % (side, i)], dtype)
        x = ops.conv2d(x, w, b, (2, 2) if i == 1 else (1, 1), (k // 2, k // 2))
        if i < 5:
            x = ops.elu(x)
        cin = cout
    return x


def stereo_forward(net, wts, left, right, dtype=torch.float32, return_intermediates=False):
    """left/right: numpy [3,H,W] or [N,3,H,W] in [0,1]  ->  disparity numpy [N?,H,W] (pixels)."""
    spec = SPECS[net]
    single = left.ndim == 3
    l = _t(left, dtype)
    r = _t(right, dtype)
    if single:
        l, r = l[None], r[None]
    inter = {}
    with torch.no_grad():
        fl = tower(wts, "left", l, spec, dtype)
        fr = tower(wts, "right", r, spec, dtype)
        x = ops.cost_volume(fl, fr, spec["max_disp"])                    # [N, D, 2C, h, w]
        inter["cost_vol"] = x
        skips = {}
        cin = x.shape[2]
        names = ("1", "2", "3ds", "4", "5", "6ds", "7", "8")
        for name, cout in zip(names, spec["enc"]):
            w = _t(wts["conv3D_%s_k" % name], dtype).view(cout, 3, cin, 3, 3)
            b = _t(wts["conv3D_%s_b" % name], dtype)
            if name.endswith("ds"):
                x = ops.pad_d(x, 1)                                       # conv3D_3ds_pad (:212)
                y = ops.conv3d(x, w, b, (2, 2, 2), (0, 1, 1))            # (:219)
            else:
                y = ops.conv3d(x, w, b, (1, 1, 1), (1, 1, 1))
            if name != "8":
                x = ops.elu(ops.transform(y))                             # T then ELU -> [N,D,C,H,W]
            else:
                x = ops.elu(y)                                            # stays [N,K,D,H,W] (:316-323)
            inter["conv3D_" + name] = x
            if name in ("2", "5"):
                skips[name] = x
            cin = cout
        # decoder (:331-415)
        for i, (cout, skip) in enumerate(zip(spec["dec"], ("5", "2", None)), start=1):
            k_in = x.shape[1]
            w = _t(wts["deconv3D_%d_k" % i], dtype).view(k_in, 3, cout, 3, 3)
            b = _t(wts["deconv3D_%d_b" % i], dtype)
            dy, hy, wy = x.shape[2:]
            out_dims = (2 * dy + 1, cout, 2 * hy - 1, 2 * wy - 1)         # e.g. {25, 64, 81, 257}
            x = ops.conv3d_transpose(x, w, b, (2, 2, 2), (0, 1, 1), out_dims)
            x = ops.slice_d(x, 0, out_dims[0] - 1)
            if skip is not None:
                x = ops.elu(x + skips[skip])
                x = ops.transform(x)                                      # [N,C,D,H,W]
            inter["deconv3D_%d" % i] = x
        disp = ops.softargmax(x, is_min=True)[:, 0]                       # [N, H, W]
    out = disp.to(torch.float32).numpy() if dtype != torch.float64 else disp.numpy()
    if single:
        out = out[0]
    if return_intermediates:
        return out, inter
    return out
