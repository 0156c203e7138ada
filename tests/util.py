import os
import struct

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def write_ref_bin(path, arr):
    """The reference's test-vector format: int32 ndims | int32 dims[] | float32 data (tests_main.cpp:259-275)."""
    arr = np.ascontiguousarray(arr, dtype=np.float32)
    with open(path, "wb") as f:
        f.write(struct.pack("<i", arr.ndim))
        f.write(struct.pack("<%di" % arr.ndim, *arr.shape))
        arr.tofile(f)


def export_fixture_dir(dst):
    from oracle.io import fixtures
    os.makedirs(dst, exist_ok=True)
    for name, arr in fixtures().items():
        write_ref_bin(os.path.join(dst, name + ".bin"), arr)
    return dst


def trailnet_model_files():
    """-> (deploy prototxt path, caffemodel path) of the TrailNet S-ResNet-18 fixtures (tests/golden/trailnet/, written by
    tests/golden/make_golden_trailnet.py).  The prototxt is stored gzip-compressed; it is unpacked once into the temp directory
    because the Caffe parser (like TensorRT's) takes file paths."""
    import gzip
    import tempfile
    tn = os.path.join(ROOT, "tests", "golden", "trailnet")
    out_dir = os.path.join(tempfile.gettempdir(), "redtail_b200_trailnet_%d" % os.getuid())
    os.makedirs(out_dir, exist_ok=True)
    proto = os.path.join(out_dir, "sresnet18_deploy.prototxt")
    text = gzip.open(os.path.join(tn, "sresnet18_deploy.prototxt.gz"), "rb").read()
    if not os.path.exists(proto) or open(proto, "rb").read() != text:
        tmp = proto + ".%d.tmp" % os.getpid()
        with open(tmp, "wb") as f:
            f.write(text)
        os.replace(tmp, proto)
    return proto, os.path.join(tn, "sresnet18_weights.caffemodel")
