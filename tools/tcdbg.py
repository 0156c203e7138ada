"""Debug helper (GPU): tcgen05 conv3d vs the CPU oracle on a few NVSmall-class shapes; prints error statistics."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from redtail_b200 import ops as R
from oracle import ops as O


def run(cin, cout, stride, transposed, prec, dims=(6, 13, 21)):
    g = torch.Generator().manual_seed(cin * 7 + cout)
    d, h, w = dims
    if not transposed:
        x = torch.randn(1, d + (stride == 2), cin, h, w, generator=g)
        if stride == 2:
            x[:, -1] = 0
        wt = torch.randn(cout, 3, cin, 3, 3, generator=g) / np.sqrt(27 * cin)
        b = torch.randn(cout, generator=g)
        pad = (1, 1, 1) if stride == 1 else (0, 1, 1)
        ref = O.conv3d(x.double(), wt.double(), b.double(), (stride,) * 3, pad).float()
        op = R.Conv3d(wt.numpy(), b.numpy(), (stride,) * 3, pad, tuple(x.shape[1:]), precision=prec)
        y = op(x.cuda())
    else:
        x = torch.randn(1, cin, d, h, w, generator=g)
        wt = torch.randn(cin, 3, cout, 3, 3, generator=g) / np.sqrt(27 * cin / 8)
        b = torch.randn(cout, generator=g)
        od = (2 * d + 1, cout, 2 * h - 1, 2 * w - 1)
        ref = O.slice_d(O.conv3d_transpose(x.double(), wt.double(), b.double(), (2, 2, 2), (0, 1, 1), od), 0, 2 * d).float()
        op = R.Conv3d(wt.numpy(), b.numpy(), (2, 2, 2), (0, 1, 1), tuple(x.shape[1:]), out_dims=od, transposed=True, precision=prec, slice_d=1)
        y = op(x.cuda())
    torch.cuda.synchronize()
    err = (y.cpu() - ref).abs()
    print("cin=%3d cout=%3d stride=%d tr=%d prec=%d dims=%s kernel=%s  max|err|=%.3e mean=%.3e  ref_max=%.2f bad=%d/%d" %
          (cin, cout, stride, transposed, prec, dims, R.last_kernel(), err.max(), err.mean(), ref.abs().max(), int((err > 1e-3).sum()), err.numel()), flush=True)


if __name__ == "__main__":
    cfgs = [(64, 32, 1, 0), (32, 32, 1, 0), (16, 16, 1, 0), (128, 128, 1, 0), (32, 64, 2, 0), (64, 128, 2, 0),
            (128, 64, 1, 1), (64, 32, 1, 1), (32, 1, 1, 1)]
    for prec in (R.PREC_FP32, R.PREC_FP16):
        for c in cfgs:
            run(*c, prec)
    run(64, 32, 1, 0, R.PREC_FP32, dims=(5, 40, 70))
    run(64, 32, 1, 1, R.PREC_FP32, dims=(3, 20, 35))
