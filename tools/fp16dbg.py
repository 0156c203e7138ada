"""Debug helper (GPU): NVSmall / NVTiny disparity error of the engine in each precision mode."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, ".")
from oracle import io as oio
for net, (h, w) in (("nvtiny", (161, 513)), ("nvsmall", (321, 1025))):
    gold = np.load("tests/golden/disp_%s_%dx%d_f64oracle.npy" % (net, w, h))
    l, r = oio.load_sample_pair(); l, r = oio.resize_pair(l, r, h, w)
    for prec in ("fp32", "fp16", "simt"):
        os.environ["REDTAIL_CONV3D_PRECISION"] = prec
        from redtail_b200 import StereoEngine
        eng = StereoEngine(net, h, w, oio.weights_path(net))
        d = eng(torch.from_numpy(l[None]).cuda(), torch.from_numpy(r[None]).cuda()); torch.cuda.synchronize()
        e = np.abs(d.cpu().numpy()[0] - gold)
        print(net, prec, "max %.3e mean %.3e p99.9 %.3e steps %d" % (e.max(), e.mean(), np.quantile(e, 0.999), eng.num_layers), flush=True)
        eng.close()
