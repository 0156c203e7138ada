#!/usr/bin/env python
"""float64 oracle disparities of NVSmall on the seeded synthetic KITTI-shaped pairs (set S2 --
oracle/io.py synthetic_pair, the generator bench.py times): tests/golden/disp_nvsmall_synth<seed>_f64oracle.npy.
Pure CPU, no reference checkout needed (the oracle and the committed weights are enough); ~2 min per seed."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

if __name__ == "__main__":
    import torch
    from oracle import nets, io as oio
    torch.set_num_threads(os.cpu_count())
    wts = oio.read_weights(oio.weights_path("nvsmall"))
    for seed in (1234, 1235, 1236):
        l, r = oio.synthetic_pair(321, 1025, seed=seed)
        d = nets.stereo_forward("nvsmall", wts, l, r, dtype=torch.float64)
        np.save(os.path.join(HERE, "disp_nvsmall_synth%d_f64oracle.npy" % seed), d.astype(np.float32))
        print(seed, float(d.min()), float(d.max()))
