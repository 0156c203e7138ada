"""GPU helper: the fused cost_vol+conv3D_1 step alone on the NVSmall shape (for ncu launch lists / captures)."""
import sys, numpy as np, torch
sys.path.insert(0, ".")
from redtail_b200 import ops
g = torch.Generator().manual_seed(1)
c, k, h, w, D = 32, 32, 161, 513, 48
l, r = torch.randn(1, c, h, w, generator=g).cuda(), torch.randn(1, c, h, w, generator=g).cuda()
wt = (torch.randn(k, 3, 2 * c, 3, 3, generator=g) / 40).numpy(); b = torch.randn(k, generator=g).numpy()
op = ops.CostVolumeConv3d(wt, b, (c, h, w), D, precision=ops.PREC_FP32, fuse_elu=True, out_layout=ops.LAYOUT_SPLIT16)
for _ in range(3): y = op(l, r)
torch.cuda.synchronize()
a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20): op(l, r)
e.record(); torch.cuda.synchronize()
print("fused cost_vol+conv3D_1: %.3f ms" % (a.elapsed_time(e) / 20))
