"""GPU helper: ResNet-18 (3-D) plan at batch B, per-layer times; used to bisect engine variants (env switches) under a timeout."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from redtail_b200 import StereoEngine
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
plan = sys.argv[2] if len(sys.argv) > 2 else "resnet18_1025x321_fp16.plan"
t0 = time.time()
with open(os.path.join("dropin", "_ref", "plans", plan), "rb") as f:
    eng = StereoEngine.deserialize(f.read(), max_batch=B)
print("engine built %.1f s" % (time.time() - t0), flush=True)
g = torch.Generator().manual_seed(0)
l = torch.rand(B, 3, 321, 1025, generator=g).cuda()
r = torch.rand(B, 3, 321, 1025, generator=g).cuda()
print("profile:", flush=True)
for name, ms in eng.profile(l, r):
    print("  %-70s %.3f ms" % (name[:70], ms), flush=True)
t0 = time.time()
for _ in range(3):
    d = eng(l, r)
torch.cuda.synchronize()
print("3 steps %.3f s, disp mean %.4f" % (time.time() - t0, float(d.mean())), flush=True)
