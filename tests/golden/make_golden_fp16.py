#!/usr/bin/env python
"""Goldens of the reference's fp16 configuration (trt_weights_fp16.bin, sample_app/main.cpp:224-256).

Runs only in the build container (needs /root/reference).  Writes under tests/golden/:

  weights/fp16_md5.json     md5 of the reference's four trt_weights_fp16.bin files.  Each of them is the elementwise fp16
                            rounding of the committed trt_weights.bin (verified here, value by value), so the tests
                            re-create the fp16 files byte-identically from weights/<net>_fp32.bin (oracle/io.py
                            write_fp16_weights) and check this md5 -- no second copy of the weights is committed.
  disp_<net>_fp16w_f64oracle.npy
                            float64 oracle disparity with the fp16 weights on the sample pair: what an exact-arithmetic
                            engine computes from trt_weights_fp16.bin.  The fp16 tolerance of north_star (1e-2 px)
                            is asserted against these.
"""
import hashlib
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/stereoDNN"
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
NETS = (("nvsmall", "NVSmall"), ("nvtiny", "NVTiny"), ("resnet18", "ResNet-18"), ("resnet18_2D", "ResNet-18_2D"))


def main():
    import torch
    from oracle import nets, io as oio, plan as oplan
    torch.set_num_threads(os.cpu_count())
    md5 = {}
    for net, d in NETS:
        ref16 = os.path.join(REF, "models", d, "TensorRT", "trt_weights_fp16.bin")
        w32 = oio.read_weights(oio.weights_path(net))
        w16 = oio.read_weights(ref16, np.float16)
        assert list(w32) == list(w16)
        for k in w32:
            assert np.array_equal(w32[k].astype(np.float16), w16[k]), (net, k)
        tmp = "/tmp/make_golden_fp16_%s.bin" % net
        oio.write_fp16_weights(oio.weights_path(net), tmp)
        with open(tmp, "rb") as f, open(ref16, "rb") as g:
            a, b = f.read(), g.read()
        assert a == b, net + ": regenerated fp16 file differs from the reference's"
        md5[net] = hashlib.md5(b).hexdigest()
    with open(os.path.join(HERE, "weights", "fp16_md5.json"), "w") as f:
        json.dump(md5, f, indent=1, sort_keys=True)
    print("fp16 weight files reproduce byte for byte:", md5)
    if "--skip-disp" in sys.argv:
        return
    left, right = oio.load_sample_pair()
    for net, (h, w) in (() if "--only-2d-1025" in sys.argv else (("nvtiny", (161, 513)), ("nvsmall", (321, 1025)))):
        wts = oio.read_weights("/tmp/make_golden_fp16_%s.bin" % net, np.float16)
        l, r = oio.resize_pair(left, right, h, w)
        disp = nets.stereo_forward(net, wts, l, r, dtype=torch.float64)
        np.save(os.path.join(HERE, "disp_%s_%dx%d_fp16w_f64oracle.npy" % (net, w, h)), disp.astype(np.float32))
        print(net, disp.shape, float(disp.min()), float(disp.max()))
    driver = os.path.join(ROOT, "dropin", "_ref", "nvstereo_net_driver")
    cases = (("resnet18_2D", (257, 513)), ("resnet18_2D", (321, 1025)), ("resnet18", (321, 1025)))
    if "--only-2d-1025" in sys.argv:
        cases = cases[1:2]
    for net, (h, w) in cases:
        tmp = "/tmp/make_golden_fp16_%s" % net
        np.zeros(3 * h * w, dtype=np.float32).tofile(tmp + ".z")
        subprocess.run([driver, net, str(w), str(h), oio.weights_path(net), tmp + ".z", tmp + ".z", tmp + ".plan", "dump"], check=True)
        with open(tmp + ".plan", "rb") as f:
            pl = oplan.parse(f.read(), round_fp16=True)
        l, r = oio.resize_pair(left, right, h, w)
        out = oplan.execute(pl, {"left": l[None].astype(np.float64), "right": r[None].astype(np.float64)})
        disp = list(out.values())[0].reshape(h, w)
        np.save(os.path.join(HERE, "disp_%s_%dx%d_fp16w_f64oracle.npy" % (net, w, h)), disp.astype(np.float32))
        print(net, disp.shape, float(disp.min()), float(disp.max()))


if __name__ == "__main__":
    main()
