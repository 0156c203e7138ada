#!/usr/bin/env python
"""Regenerates the committed golden fixtures from the read-only reference checkout.

Runs only in the build container (needs /root/reference); the GPU box uses the
committed outputs.  What it writes (all under tests/golden/):

  plugin_fixtures.npz       the 58 known-answer tensors of the reference's plugin
                            unit tests (stereoDNN/tests/data/*.bin, produced by
                            stereoDNN/scripts/test_data_generator.py with TF ops),
                            re-packed into one npz keyed by file stem.
  weights/<net>_fp32.bin    the reference's trained weights, byte-identical copies of
                            stereoDNN/models/<net>/TensorRT/trt_weights.bin (data, not
                            source; format: cstring name, u32 count, count x f32 --
                            writer stereoDNN/scripts/tensorrt_model_builder.py:52-60).
  images/kitti_{left,right}_1025x321.f16.npy
                            the sample stereo pair of stereoDNN/sample_app/data
                            (img_{left,right}.bin, CHW float32 in [0,1]) stored as fp16.
  disp_*.npy                network-level golden disparities (NVSmall, NVTiny: oracle/nets.py; ResNet-18 nets: the reference's
                            generated builders dumped as plans and executed by oracle/plan.py) computed by the fixture-pinned
                            CPU oracle (oracle/nets.py, float64) on that pair.

Usage: python tests/golden/make_golden.py [--skip-disp]
"""
import os
import shutil
import struct
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/stereoDNN"
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def read_bin(path):
    """int32 ndims | int32 dims[] | float32 data  (reference reader: tests/tests_main.cpp:259-275)."""
    with open(path, "rb") as f:
        raw = f.read()
    nd = struct.unpack_from("<i", raw, 0)[0]
    dims = struct.unpack_from("<%di" % nd, raw, 4)
    data = np.frombuffer(raw, dtype="<f4", offset=4 + 4 * nd)
    assert data.size == int(np.prod(dims)), path
    return data.reshape(dims).copy()


def main():
    data_dir = os.path.join(REF, "tests", "data")
    fx = {}
    for name in sorted(os.listdir(data_dir)):
        if name.endswith(".bin"):
            fx[name[:-4]] = read_bin(os.path.join(data_dir, name))
    assert len(fx) == 58, len(fx)
    np.savez_compressed(os.path.join(HERE, "plugin_fixtures.npz"), **fx)
    print("plugin_fixtures.npz:", len(fx), "tensors")

    os.makedirs(os.path.join(HERE, "weights"), exist_ok=True)
    for net, d in (("nvsmall", "NVSmall"), ("nvtiny", "NVTiny"), ("resnet18", "ResNet-18"), ("resnet18_2D", "ResNet-18_2D")):
        shutil.copyfile(os.path.join(REF, "models", d, "TensorRT", "trt_weights.bin"),
                        os.path.join(HERE, "weights", net + "_fp32.bin"))
        os.chmod(os.path.join(HERE, "weights", net + "_fp32.bin"), 0o644)

    os.makedirs(os.path.join(HERE, "images"), exist_ok=True)
    for side in ("left", "right"):
        img = np.fromfile(os.path.join(REF, "sample_app", "data", "img_%s.bin" % side), dtype="<f4")
        img = img.reshape(3, 321, 1025)
        np.save(os.path.join(HERE, "images", "kitti_%s_1025x321.f16.npy" % side), img.astype(np.float16))

    if "--skip-disp" in sys.argv:
        return
    import torch
    from oracle import nets, io as oio
    torch.set_num_threads(os.cpu_count())
    left, right = oio.load_sample_pair()
    for net, (h, w) in (("nvtiny", (161, 513)), ("nvsmall", (321, 1025))):
        wts = oio.read_weights(os.path.join(HERE, "weights", net + "_fp32.bin"))
        l, r = oio.resize_pair(left, right, h, w)
        disp = nets.stereo_forward(net, wts, l, r, dtype=torch.float64)
        np.save(os.path.join(HERE, "disp_%s_%dx%d_f64oracle.npy" % (net, w, h)), disp.astype(np.float32))
        print(net, disp.shape, float(disp.min()), float(disp.max()))
    # ResNet-18 (3-D, 1025x321) and ResNet18_2D (513x257): the reference's own generated builders
    # (sample_app/resnet18_*_net.cpp, 1036 / 777 lines) are the wiring.  tools/dropin/build.sh compiles them unchanged against
    # include/NvInfer.h; the net driver's host-only `dump` mode writes the network plan (no GPU), and oracle/plan.py executes
    # that plan with the fixture-pinned ops in float64.
    import subprocess
    from oracle import plan as oplan
    driver = os.path.join(os.path.dirname(os.path.dirname(HERE)), "dropin", "_ref", "nvstereo_net_driver")
    for net, (h, w) in (("resnet18_2D", (257, 513)), ("resnet18", (321, 1025))):
        tmp = "/tmp/make_golden_%s" % net
        np.zeros(3 * h * w, dtype=np.float32).tofile(tmp + ".z")
        subprocess.run([driver, net, str(w), str(h), os.path.join(HERE, "weights", net + "_fp32.bin"), tmp + ".z", tmp + ".z",
                        tmp + ".plan", "dump"], check=True)
        with open(tmp + ".plan", "rb") as f:
            pl = oplan.parse(f.read())
        l, r = oio.resize_pair(left, right, h, w)
        out = oplan.execute(pl, {"left": l[None].astype(np.float64), "right": r[None].astype(np.float64)})
        disp = list(out.values())[0].reshape(h, w)
        np.save(os.path.join(HERE, "disp_%s_%dx%d_f64oracle.npy" % (net, w, h)), disp.astype(np.float32))
        print(net, disp.shape, float(disp.min()), float(disp.max()))


if __name__ == "__main__":
    main()
