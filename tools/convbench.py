"""GPU helper: NVSmall-class 3-D conv layers (split16 in/out, fused ELU) timed alone for a list of plan variants, all in ONE
process (a fresh process costs seconds of CUDA start-up on the box).
   python tools/convbench.py '{"ENV": "v", ...}' ...        (the default plan is always measured first)"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from redtail_b200 import ops

LAYERS = [("conv3D_2", 32, 32, 48, 161, 513, 1), ("conv3D_3ds", 32, 64, 48, 161, 513, 2), ("conv3D_4", 64, 64, 24, 81, 257, 1),
          ("conv3D_6ds", 64, 128, 24, 81, 257, 2), ("conv3D_7", 128, 128, 12, 41, 129, 1)]
if os.environ.get("CONVBENCH_LAYERS"):
    LAYERS = [l for l in LAYERS if l[0] in os.environ["CONVBENCH_LAYERS"].split(",")]
variants = [{}] + [json.loads(a) for a in sys.argv[1:]]
data = {}
for name, cin, cout, d, h, w, stride in LAYERS:
    g = torch.Generator().manual_seed(1)
    x = torch.randn(1, d, cin, h, w, generator=g).cuda()
    wt = (torch.randn(cout, 3, cin, 3, 3, generator=g) / np.sqrt(27 * cin)).numpy()
    b = torch.randn(cout, generator=g).numpy()
    data[name] = (ops.dense_to_split16(x), wt, b)
    del x
for env in variants:
    line = []
    for name, cin, cout, d, h, w, stride in LAYERS:
        xs, wt, b = data[name]
        pad = (1, 1, 1) if stride == 1 else (0, 1, 1)
        os.environ.update(env)
        try:
            op = ops.Conv3d(wt, b, (stride,) * 3, pad, (d, cin, h, w), precision=ops.PREC_FP32, fuse_elu=True,
                            in_layout=ops.LAYOUT_SPLIT16, out_layout=ops.LAYOUT_SPLIT16, pad_end_d=1 if stride == 2 else 0)
        except ops.RedtailError as e:
            line.append("%s: %s" % (name, e))
            continue
        finally:
            for k in env:
                os.environ.pop(k)
        for _ in range(3):
            y = op(xs)
        torch.cuda.synchronize()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            op(xs)
        e.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(e) / 10
        od = op.out_dims
        fl = 2.0 * cout * 27 * cin * od[1] * od[2] * od[3]
        line.append("%s %.3f ms (%.0f TF/s)" % (name, ms, fl / ms / 1e9))
        del op, y
    print(json.dumps(env), " | ".join(line), flush=True)
