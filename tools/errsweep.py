"""GPU helper: NVSmall disparity error vs the float64 oracle golden for a few engine variants (max / mean / tail / where)."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, ".")
from oracle import io as oio
from redtail_b200 import StereoEngine

gold = np.load(os.path.join(oio.GOLDEN, "disp_nvsmall_1025x321_f64oracle.npy"))
l, r = oio.load_sample_pair(); l, r = oio.resize_pair(l, r, 321, 1025)
lt, rt = torch.from_numpy(l[None]).cuda(), torch.from_numpy(r[None]).cuda()
variants = [{}] + [json.loads(a) for a in sys.argv[1:]]
prev = None
for env in variants:
    os.environ.update(env)
    eng = StereoEngine("nvsmall", 321, 1025, oio.weights_path("nvsmall"))
    for k in env: os.environ.pop(k)
    d = eng(lt, rt).cpu().numpy()[0]
    err = np.abs(d - gold)
    iy, ix = np.unravel_index(err.argmax(), err.shape)
    top = np.sort(err.ravel())[-5:][::-1]
    print(env, "max %.3g at (%d,%d) mean %.3g p99.99 %.3g top5 %s  n>5e-4: %d" % (err.max(), iy, ix, err.mean(), np.quantile(err, 0.9999),
          np.array2string(top, precision=2), (err > 5e-4).sum()), flush=True)
    del eng
