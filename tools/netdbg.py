"""Debug helper (GPU): NVSmall 3-D stack layer by layer, tcgen05 (fp16x2 split) vs fp32 SIMT, same inputs per layer."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from redtail_b200 import ops as R
from oracle import io as oio

net = sys.argv[1] if len(sys.argv) > 1 else "nvsmall"
H, W, D = (321, 1025, 48) if net == "nvsmall" else (161, 513, 24)
wts = oio.read_weights(oio.weights_path(net))
l, r = oio.load_sample_pair()
l, r = oio.resize_pair(l, r, H, W)


def tower(x, side):
    cin = 3
    for i in range(1, 6):
        k = 5 if i == 1 else 3
        b = wts["%s_conv%d_b" % (side, i)]
        w = wts["%s_conv%d_k" % (side, i)].reshape(len(b), cin, k, k)
        op = R.Conv2d(w, b, (2, 2) if i == 1 else (1, 1), (k // 2, k // 2), tuple(x.shape[2:]), fuse_elu=i < 5)
        x = op(x)
        cin = len(b)
    return x


fl = tower(torch.from_numpy(l[None]).cuda(), "left")
fr = tower(torch.from_numpy(r[None]).cuda(), "right")
x = R.cost_volume(fl, fr, D)
print("cost_vol", tuple(x.shape), "absmax %.3f" % x.abs().max().item())
skips = {}
x_tc = x
cin = x.shape[2]
for name in ("1", "2", "3ds", "4", "5", "6ds", "7", "8"):
    b = wts["conv3D_%s_b" % name]
    w = wts["conv3D_%s_k" % name].reshape(len(b), 3, cin, 3, 3)
    ds = name.endswith("ds")
    if ds:
        x = R.pad_d(x, 1)
        x_tc = R.pad_d(x_tc, 1)
    st, pad = ((2, 2, 2), (0, 1, 1)) if ds else ((1, 1, 1), (1, 1, 1))
    kw = dict(fuse_elu=True, out_transposed=name != "8")
    op_s = R.Conv3d(w, b, st, pad, tuple(x.shape[1:]), precision=R.PREC_SIMT, **kw)
    op_t = R.Conv3d(w, b, st, pad, tuple(x.shape[1:]), precision=R.PREC_FP32, **kw)
    y_s = op_s(x)
    y_t_same = op_t(x)          # TC on the SIMT chain's input: per-layer error
    y_t_chain = op_t(x_tc)      # TC chain
    e1 = (y_t_same - y_s).abs()
    e2 = (y_t_chain - y_s).abs()
    print("conv3D_%-3s out %s absmax %.2f | wmax %.3f | per-layer max %.2e mean %.2e | chain max %.2e mean %.2e" %
          (name, tuple(y_s.shape), y_s.abs().max().item(), np.abs(w).max(), e1.max().item(), e1.mean().item(), e2.max().item(), e2.mean().item()), flush=True)
    x, x_tc = y_s, y_t_chain
    if name in ("2", "5"):
        skips[name] = (y_s, y_t_chain)
    cin = len(b)
    del op_s, op_t, y_t_same
for i, sk in zip((1, 2, 3), ("5", "2", None)):
    b = wts["deconv3D_%d_b" % i]
    kin = x.shape[1]
    w = wts["deconv3D_%d_k" % i].reshape(kin, 3, len(b), 3, 3)
    dy, hy, wy = x.shape[2:]
    od = (2 * dy + 1, len(b), 2 * hy - 1, 2 * wy - 1)
    kw = dict(out_dims=od, transposed=True, slice_d=1, fuse_elu=sk is not None)
    op_s = R.Conv3d(w, b, (2, 2, 2), (0, 1, 1), tuple(x.shape[1:]), precision=R.PREC_SIMT, **kw)
    op_t = R.Conv3d(w, b, (2, 2, 2), (0, 1, 1), tuple(x.shape[1:]), precision=R.PREC_FP32, **kw)
    s_s, s_t = (skips[sk] if sk else (None, None))
    y_s = op_s(x, s_s)
    y_t_same = op_t(x, s_s)
    y_t_chain = op_t(x_tc, s_t)
    e1 = (y_t_same - y_s).abs()
    e2 = (y_t_chain - y_s).abs()
    print("deconv3D_%d out %s absmax %.2f | per-layer max %.2e mean %.2e | chain max %.2e mean %.2e" %
          (i, tuple(y_s.shape), y_s.abs().max().item(), e1.max().item(), e1.mean().item(), e2.max().item(), e2.mean().item()), flush=True)
    if sk:
        y_s, y_t_chain = R.transform(y_s), R.transform(y_t_chain)
    x, x_tc = y_s, y_t_chain
    del op_s, op_t, y_t_same
d_s = R.softargmax(x, True)
d_t = R.softargmax(x_tc, True)
gold = np.load("tests/golden/disp_%s_%dx%d_f64oracle.npy" % (net, W, H))
print("disp: TC-vs-SIMT max %.2e | SIMT-vs-oracle max %.2e | TC-vs-oracle max %.2e" %
      ((d_s - d_t).abs().max().item(), np.abs(d_s.cpu().numpy()[0, 0] - gold).max(), np.abs(d_t.cpu().numpy()[0, 0] - gold).max()))
e = np.abs(d_t.cpu().numpy()[0, 0] - gold)
iy, ix = np.unravel_index(e.argmax(), e.shape)
print("worst pixel", iy, ix, "logit spread at worst:", x[0, :, 0, iy, ix].min().item(), x[0, :, 0, iy, ix].max().item())
