"""GPU helper: runs the reference's unchanged ResNet-18 builders through this repo's engine (tools/dropin net driver) on the
sample pair and stores the raw outputs under gpurun_out/ for comparison with the plan-oracle goldens."""
import os, subprocess, sys
import numpy as np
sys.path.insert(0, ".")
from oracle import io as oio
exe = "dropin/_ref/nvstereo_net_driver"
os.makedirs("gpurun_out", exist_ok=True)
l0, r0 = oio.load_sample_pair()
for net, (h, w) in (("resnet18_2D", (257, 513)), ("resnet18", (321, 1025))):
    l, r = oio.resize_pair(l0, r0, h, w)
    l.tofile("/tmp/l.bin"); r.tofile("/tmp/r.bin")
    out = "gpurun_out/disp_%s_engine.bin" % net
    p = subprocess.run([exe, net, str(w), str(h), oio.weights_path(net), "/tmp/l.bin", "/tmp/r.bin", out, "profile"],
                       capture_output=True, text=True, timeout=600)
    print(net, "rc", p.returncode)
    print(p.stdout[-1500:])
    print(p.stderr[-1500:])
