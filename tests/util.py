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
