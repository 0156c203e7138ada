"""GPU helper: NVSmall disparity error vs the float64 oracle goldens for a few engine variants (max / mean / tail / where).
   python tools/errsweep.py <pair> '{"ENV": "v"}' ...     pair = kitti | synth1234 | synth1235 | synth1236 | all"""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, ".")
from oracle import io as oio
from redtail_b200 import StereoEngine


def pair(name):
    if name == "kitti":
        l, r = oio.load_sample_pair(); l, r = oio.resize_pair(l, r, 321, 1025)
        gold = np.load(os.path.join(oio.GOLDEN, "disp_nvsmall_1025x321_f64oracle.npy"))
    else:
        seed = int(name[5:])
        l, r = oio.synthetic_pair(321, 1025, seed=seed)
        gold = np.load(os.path.join(oio.GOLDEN, "disp_nvsmall_synth%d_f64oracle.npy" % seed))
    return torch.from_numpy(l[None]).cuda(), torch.from_numpy(r[None]).cuda(), gold


names = ["kitti", "synth1234", "synth1235", "synth1236"] if sys.argv[1] == "all" else [sys.argv[1]]
pairs = {n: pair(n) for n in names}
variants = [{}] + [json.loads(a) for a in sys.argv[2:]]
for env in variants:
    os.environ.update(env)
    eng = StereoEngine("nvsmall", 321, 1025, oio.weights_path("nvsmall"))
    for k in env: os.environ.pop(k)
    for n, (lt, rt, gold) in pairs.items():
        d = eng(lt, rt).cpu().numpy()[0]
        err = np.abs(d - gold)
        iy, ix = np.unravel_index(err.argmax(), err.shape)
        top = np.sort(err.ravel())[-5:][::-1]
        print(env, n, "max %.3g at (%d,%d) mean %.3g p99.99 %.3g top5 %s  n>5e-4: %d n>1e-3: %d" % (err.max(), iy, ix, err.mean(), np.quantile(err, 0.9999),
              np.array2string(top, precision=2), (err > 5e-4).sum(), (err > 1e-3).sum()), flush=True)
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    lt, rt, _ = pairs[names[0]]
    for _ in range(3): eng(lt, rt)
    t0.record()
    for _ in range(10): eng(lt, rt)
    t1.record(); torch.cuda.synchronize()
    print(env, "ms/pair %.3f" % (t0.elapsed_time(t1) / 10), flush=True)
    del eng
