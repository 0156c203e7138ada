"""CPU: rt_write_png16 (host code of libredtail_b200.so, no device needed) writes what cv::imwrite writes for a CV_16U Mat --
a 16-bit greyscale PNG that standard decoders read back bit-exactly (sample_app/main.cpp:317-330)."""
import struct
import zlib

import numpy as np


def _decode_png16(path):
    """Minimal PNG reader (zlib from the standard library): 16-bit greyscale, filter type 0 only."""
    raw = open(path, "rb").read()
    assert raw[:8] == b"\x89PNG\r\n\x1a\n"
    pos, idat, hdr = 8, b"", None
    while pos < len(raw):
        n, typ = struct.unpack(">I4s", raw[pos:pos + 8])
        data = raw[pos + 8:pos + 8 + n]
        crc = struct.unpack(">I", raw[pos + 8 + n:pos + 12 + n])[0]
        assert zlib.crc32(typ + data) & 0xFFFFFFFF == crc, typ
        if typ == b"IHDR":
            hdr = struct.unpack(">IIBBBBB", data)
        elif typ == b"IDAT":
            idat += data
        pos += 12 + n
    w, h, depth, ctype, comp, flt, inter = hdr
    assert (depth, ctype, comp, flt, inter) == (16, 0, 0, 0, 0)
    lines = zlib.decompress(idat)
    assert len(lines) == h * (1 + 2 * w)
    out = np.zeros((h, w), np.uint16)
    for y in range(h):
        row = lines[y * (1 + 2 * w):(y + 1) * (1 + 2 * w)]
        assert row[0] == 0
        out[y] = np.frombuffer(row[1:], dtype=">u2")
    return out


def test_png16_roundtrip(tmp_path):
    from redtail_b200 import ops
    rng = np.random.default_rng(3)
    for h, w in ((1, 1), (7, 13), (321, 1025)):          # the last one spans several 64 KB stored blocks
        a = rng.integers(0, 65536, (h, w)).astype(np.uint16)
        p = tmp_path / ("d_%dx%d.png" % (w, h))
        ops.write_png16(p, a)
        assert np.array_equal(_decode_png16(p), a)
    try:
        import cv2
    except ImportError:
        return
    b = cv2.imread(str(tmp_path / "d_1025x321.png"), cv2.IMREAD_UNCHANGED)
    assert b.dtype == np.uint16 and np.array_equal(b, a)
