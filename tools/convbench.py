"""GPU helper: one NVSmall-class 3-D conv layer (split16 in/out, fused ELU) timed alone; args: cin cout d h w [stride];
CONVBENCH_PREC=fp16 selects the single-product mode (one accumulation chain per tile)."""
import os
import sys, numpy as np, torch
sys.path.insert(0, ".")
from redtail_b200 import ops
cin, cout, d, h, w = [int(a) for a in sys.argv[1:6]]
stride = int(sys.argv[6]) if len(sys.argv) > 6 else 1
g = torch.Generator().manual_seed(1)
x = torch.randn(1, d, cin, h, w, generator=g).cuda()
wt = (torch.randn(cout, 3, cin, 3, 3, generator=g) / np.sqrt(27 * cin)).numpy(); b = torch.randn(cout, generator=g).numpy()
pad = (1, 1, 1) if stride == 1 else (0, 1, 1)
op = ops.Conv3d(wt, b, (stride,) * 3, pad, (d, cin, h, w), precision=ops.PREC_FP16 if os.environ.get('CONVBENCH_PREC') == 'fp16' else ops.PREC_FP32, fuse_elu=True,
                in_layout=ops.LAYOUT_SPLIT16, out_layout=ops.LAYOUT_SPLIT16, pad_end_d=1 if stride == 2 else 0)
xs = ops.dense_to_split16(x)
for _ in range(3): y = op(xs)
torch.cuda.synchronize()
a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10): op(xs)
e.record(); torch.cuda.synchronize()
ms = a.elapsed_time(e) / 10
od = op.out_dims
fl = 2.0 * cout * 27 * cin * od[1] * od[2] * od[3]
print("conv %d->%d %dx%dx%d s%d: %.3f ms  %.1f TFLOP/s (algorithmic)" % (cin, cout, d, h, w, stride, ms, fl / ms / 1e9))
