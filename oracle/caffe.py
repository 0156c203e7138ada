"""TEST INFRASTRUCTURE (CPU oracle).  Caffe model reader + float64 executor for the TrailNet S-ResNet-18 classifier.

Restates what the reference runs through TensorRT's Caffe parser (ros/packages/caffe_ros/src/tensor_net.cpp:79-124:
`parser->parse(prototxt, caffemodel, network, dtype)`, input preprocessing :303-336) for
models/pretrained/TrailNet_SResNet-18.{prototxt,caffemodel}: Scale, Convolution, ReLU, Pooling (Caffe's ceil-mode output
size), Eltwise SUM, InnerProduct, Softmax, Concat -- the layer semantics are Caffe's (BVLC caffe 1.0, the framework the model
was trained with; not vendored in /root/reference, its published layer definitions are restated here).

PINNED: `tests/test_oracle_trailnet.py` runs this executor on the reference's five test images
(ros/packages/caffe_ros/tests/data/*.jpg, preprocessed exactly as tensor_net.cpp:303-336 does, fixtures in tests/golden/trailnet/)
and reproduces the six softmax outputs the reference's own test expects (ros/packages/caffe_ros/tests/tests.cpp:64-69) within
its tolerance of 1e-3.

Only tests/, __graft_entry__.smoke() and bench.py's CPU baseline may import this module.
"""
import re
import struct

import numpy as np


# ---------------------------------------------------------------------------------------------------------------------
# prototxt (protobuf text format): a small recursive-descent parser -> nested dicts, repeated fields as lists
# ---------------------------------------------------------------------------------------------------------------------
_TOKEN = re.compile(r'\s*(?:#[^\n]*\n\s*)*("(?:[^"\\]|\\.)*"|[{}:]|[^\s{}:"]+)')


def parse_prototxt(text):
    toks = _TOKEN.findall(text)
    pos = 0

    def message(top):
        nonlocal pos
        out = {}
        while pos < len(toks):
            t = toks[pos]
            if t == "}":
                if top:
                    raise ValueError("unbalanced '}' in prototxt")
                pos += 1
                return out
            name = t
            pos += 1
            if toks[pos] == ":":
                pos += 1
                if toks[pos] == "{":                       # "name: { ... }" form
                    pos += 1
                    val = message(False)
                else:
                    val = _scalar(toks[pos])
                    pos += 1
            elif toks[pos] == "{":
                pos += 1
                val = message(False)
            else:
                raise ValueError("prototxt: expected ':' or '{' after %r" % name)
            out.setdefault(name, []).append(val)
        if not top:
            raise ValueError("prototxt: missing '}'")
        return out

    return message(True)


def _scalar(tok):
    if tok.startswith('"'):
        return tok[1:-1]
    try:
        return int(tok)
    except ValueError:
        pass
    try:
        return float(tok)
    except ValueError:
        return tok                                          # enum / bool identifier


def _one(msg, key, default=None):
    v = msg.get(key)
    return v[0] if v else default


# ---------------------------------------------------------------------------------------------------------------------
# caffemodel (binary protobuf, caffe.proto): NetParameter.layer (field 100) -> name (1), type (2), blobs (7);
# BlobProto: shape (7) { dim (1, packed int64) }, data (5, packed float), legacy num/channels/height/width (1..4)
# ---------------------------------------------------------------------------------------------------------------------
def _varint(buf, pos):
    shift, val = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        val |= (b & 0x7F) << shift
        if not b & 0x80:
            return val, pos
        shift += 7


def _fields(buf, start, end):
    pos = start
    while pos < end:
        key, pos = _varint(buf, pos)
        fno, wt = key >> 3, key & 7
        if wt == 0:
            val, pos = _varint(buf, pos)
            yield fno, wt, val, None
        elif wt == 1:
            yield fno, wt, None, (pos, pos + 8)
            pos += 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            yield fno, wt, None, (pos, pos + ln)
            pos += ln
        elif wt == 5:
            yield fno, wt, None, (pos, pos + 4)
            pos += 4
        else:
            raise ValueError("caffemodel: unsupported wire type %d" % wt)


def _blob(buf, start, end):
    dims, legacy, data = None, {}, None
    floats = []
    for fno, wt, val, span in _fields(buf, start, end):
        if fno == 7 and wt == 2:                            # BlobShape
            dims = []
            for f2, w2, v2, s2 in _fields(buf, span[0], span[1]):
                if f2 == 1 and w2 == 2:                     # packed
                    p = s2[0]
                    while p < s2[1]:
                        v, p = _varint(buf, p)
                        dims.append(v)
                elif f2 == 1 and w2 == 0:
                    dims.append(v2)
        elif fno == 5 and wt == 2:                          # packed float data
            data = np.frombuffer(buf, dtype="<f4", count=(span[1] - span[0]) // 4, offset=span[0])
        elif fno == 5 and wt == 5:
            floats.append(struct.unpack_from("<f", buf, span[0])[0])
        elif fno in (1, 2, 3, 4) and wt == 0:
            legacy[fno] = val
    if data is None:
        data = np.asarray(floats, dtype=np.float32)
    if dims is None:
        dims = [legacy.get(i, 1) for i in (1, 2, 3, 4)] if legacy else [data.size]
    return np.array(data, dtype=np.float32).reshape(dims)


def read_caffemodel(path):
    """-> {layer name: [blob arrays]} for every layer that carries blobs."""
    with open(path, "rb") as f:
        buf = f.read()
    out = {}
    for fno, wt, val, span in _fields(buf, 0, len(buf)):
        if fno != 100 or wt != 2:                           # NetParameter.layer (V2 LayerParameter)
            continue
        name, blobs = None, []
        for f2, w2, v2, s2 in _fields(buf, span[0], span[1]):
            if f2 == 1 and w2 == 2:
                name = buf[s2[0]:s2[1]].decode()
            elif f2 == 7 and w2 == 2:
                blobs.append(_blob(buf, s2[0], s2[1]))
        if name is not None and blobs:
            out[name] = blobs
    return out


# ---------------------------------------------------------------------------------------------------------------------
# executor (float64, plain numpy / torch-free so that the semantics are explicit)
# ---------------------------------------------------------------------------------------------------------------------
def _conv2d(x, w, b, stride, pad):
    import torch
    import torch.nn.functional as F
    y = F.conv2d(torch.from_numpy(x), torch.from_numpy(w), None if b is None else torch.from_numpy(b), stride=stride, padding=pad)
    return y.numpy()


def _pool(x, kind, k, stride, pad):
    """Caffe PoolingLayer: output = ceil((in + 2 pad - k) / stride) + 1, clipped so that the last window starts inside the
    (padded) image; MAX ignores the padding, AVE divides by the window area clipped to the padded image (pooling_layer.cpp)."""
    n, c, h, w = x.shape
    oh = int(np.ceil((h + 2 * pad - k) / stride)) + 1
    ow = int(np.ceil((w + 2 * pad - k) / stride)) + 1
    if pad > 0:
        if (oh - 1) * stride >= h + pad:
            oh -= 1
        if (ow - 1) * stride >= w + pad:
            ow -= 1
    y = np.empty((n, c, oh, ow), x.dtype)
    for i in range(oh):
        hs, he = i * stride - pad, min(i * stride - pad + k, h + pad)
        for j in range(ow):
            ws, we = j * stride - pad, min(j * stride - pad + k, w + pad)
            area = (he - hs) * (we - ws)
            win = x[:, :, max(hs, 0):min(he, h), max(ws, 0):min(we, w)]
            y[:, :, i, j] = win.max(axis=(2, 3)) if kind == "MAX" else win.sum(axis=(2, 3)) / area
    return y


def run_net(prototxt_text, blobs, data, dtype=np.float64, return_all=False):
    """Executes the layers of a deploy prototxt in file order.  data: [N,C,H,W].  -> output of the last layer (or every blob)."""
    net = parse_prototxt(prototxt_text)
    env = {_one(net, "input", "data"): np.asarray(data, dtype)}
    last = None
    for layer in net.get("layer", []):
        name, typ = _one(layer, "name"), _one(layer, "type")
        bottoms = [env[b] for b in layer.get("bottom", [])]
        wts = [np.asarray(b, dtype) for b in blobs.get(name, [])]
        if typ == "Scale":
            sp = _one(layer, "scale_param", {})
            scale = wts[0] if wts else np.full(bottoms[0].shape[1], _one(_one(sp, "filler", {}), "value", 1.0), dtype)
            y = bottoms[0] * scale.reshape(1, -1, 1, 1)
            if _one(sp, "bias_term", "false") == "true":
                bias = wts[1] if len(wts) > 1 else np.full(bottoms[0].shape[1], _one(_one(sp, "bias_filler", {}), "value", 0.0), dtype)
                y = y + bias.reshape(1, -1, 1, 1)
        elif typ == "Convolution":
            cp = _one(layer, "convolution_param")
            y = _conv2d(bottoms[0], wts[0], wts[1] if len(wts) > 1 else None, _one(cp, "stride", 1), _one(cp, "pad", 0))
        elif typ == "ReLU":
            y = np.maximum(bottoms[0], 0)
        elif typ == "Pooling":
            pp = _one(layer, "pooling_param")
            y = _pool(bottoms[0], _one(pp, "pool", "MAX"), _one(pp, "kernel_size"), _one(pp, "stride", 1), _one(pp, "pad", 0))
        elif typ == "Eltwise":
            y = bottoms[0] + bottoms[1]                      # operation defaults to SUM (the only one the model uses)
            assert _one(_one(layer, "eltwise_param", {}), "operation", "SUM") == "SUM"
        elif typ == "InnerProduct":
            x2 = bottoms[0].reshape(bottoms[0].shape[0], -1)
            y = x2 @ wts[0].reshape(wts[0].shape[0], -1).T
            if len(wts) > 1:
                y = y + wts[1].reshape(1, -1)
        elif typ == "Softmax":
            z = bottoms[0] - bottoms[0].max(axis=1, keepdims=True)
            e = np.exp(z)
            y = e / e.sum(axis=1, keepdims=True)
        elif typ == "Concat":
            y = np.concatenate(bottoms, axis=_one(_one(layer, "concat_param", {}), "axis", 1))
        else:
            raise ValueError("caffe oracle: unsupported layer type %s (%s)" % (typ, name))
        env[_one(layer, "top")] = y
        last = y
    return env if return_all else last


def preprocess_bgr8(img_bgr8, dst_w, dst_h, scale=1.0, shift=0.0):
    """tensor_net.cpp:303-336 for a bgr8 image and InputFormat::BGR: float conversion, anisotropic cv::resize INTER_CUBIC,
    scale, shift, HWC -> CHW.  Needs cv2 (only where the fixtures are generated)."""
    import cv2
    img = img_bgr8.astype(np.float32)
    img = cv2.resize(img, (dst_w, dst_h), interpolation=cv2.INTER_CUBIC)
    if scale != 1:
        img = img * np.float32(scale)
    if shift != 0:
        img = img + np.float32(shift)
    return np.ascontiguousarray(img.transpose(2, 0, 1))
