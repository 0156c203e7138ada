"""The reference's weight-file format (writer stereoDNN/scripts/tensorrt_model_builder.py:52-60, reader
sample_app/main.cpp:111-134): a sequence of  cstring name | u32 count | count x (f32 | f16)  entries; shapes are not stored."""
import struct

import numpy as np


def read_weight_file(path, dtype=np.float32):
    """-> dict name -> flat numpy array, in file order."""
    out = {}
    esz = np.dtype(dtype).itemsize
    with open(path, "rb") as f:
        raw = f.read()
    i = 0
    while i < len(raw):
        j = raw.index(b"\0", i)
        (cnt,) = struct.unpack_from("<I", raw, j + 1)
        out[raw[i:j].decode()] = np.frombuffer(raw, dtype=dtype, count=cnt, offset=j + 5).copy()
        i = j + 5 + cnt * esz
    return out


def write_fp16_weight_file(src_fp32, dst):
    """trt_weights.bin -> trt_weights_fp16.bin: same entries, payloads rounded to fp16 (what the reference's generator writes
    next to the fp32 file; byte-identical to the reference's files for all four nets, tests/golden/make_golden_fp16.py)."""
    with open(dst, "wb") as f:
        for name, a in read_weight_file(src_fp32).items():
            f.write(name.encode() + b"\0" + struct.pack("<I", a.size) + a.astype("<f2").tobytes())
    return dst
