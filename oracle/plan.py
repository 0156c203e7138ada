"""Graph-level oracle: executes a network PLAN with the fixture-pinned CPU ops of oracle/ops.py.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): nothing in the product imports this.

A plan is what ICudaEngine::serialize() / the host-only network dump write (the engine's EngineImpl::serializeNetwork,
format documented in DESIGN.md 1a): the layer list of an INetworkDefinition with every parameter, weight and plugin blob.
Dumped host-side (no GPU) from the reference's UNCHANGED generated builders -- stereoDNN/sample_app/*_net.cpp, compiled
against include/NvInfer.h by tools/dropin -- it is the reference's own wiring of NVSmall, NVTiny, ResNet-18 and
ResNet18_2D; interpreting it with the oracle ops gives the network-level reference without restating 1000-line
builders by hand.  Pinned on CPU by reproducing oracle/nets.py's NVTiny result to 2e-6 px and the slab-wise 3-D convolutions against the
plain ops (tests/test_oracle_plan.py).

Byte formats: engine.cpp (PlanWriter) for the container, plugins.cpp (blob()) for the plugin payloads; the first three
plugin tags and their payloads are the reference's (lib/elu_plugin.cpp:170-190, cost_volume_plugin.cpp:141-170,
softargmax_plugin.cpp:207-230).
"""
import struct

import numpy as np
import torch

from . import ops

MAGIC = b"RTB2PLAN"
MAX_DIMS = 8
K_CONV, K_DECONV, K_SCALE, K_ELTWISE, K_CONCAT, K_ACTIVATION, K_SHUFFLE, K_PLUGIN, K_POOLING, K_FC, K_SOFTMAX = range(11)
P_ELU, P_COSTVOL, P_SOFTARGMAX, P_CONV3D, P_CONV3D_T, P_TRANSFORM, P_PADDING, P_SLICE = range(8)


def _f16(a):
    """Values rounded to fp16 (what the reference's trt_weights_fp16.bin holds: it is the elementwise fp16 rounding of
    trt_weights.bin, checked byte for byte by tests/golden/make_golden_fp16.py)."""
    return a.astype(np.float16).astype(np.float64)


class _Reader:
    def __init__(self, buf, pos=0, round_fp16=False):
        self.b, self.p, self.round_fp16 = buf, pos, round_fp16

    def get(self, fmt):
        v = struct.unpack_from("<" + fmt, self.b, self.p)
        self.p += struct.calcsize("<" + fmt)
        return v[0] if len(v) == 1 else v

    def str(self):
        n = self.get("i")
        s = self.b[self.p:self.p + n].decode()
        self.p += n
        return s

    def dims(self):
        nb = self.get("i")
        d = self.get("%di" % MAX_DIMS)
        return tuple(d[:nb])

    def align8(self):
        self.p = (self.p + 7) & ~7

    def weights(self):
        typ, count = self.get("i"), self.get("q")
        self.align8()
        dt = np.float16 if typ == 1 else np.float32
        a = np.frombuffer(self.b, dtype=dt, count=count, offset=self.p).astype(np.float64) if count else None
        self.p += count * np.dtype(dt).itemsize
        return _f16(a) if (self.round_fp16 and a is not None) else a


def parse(buf, round_fp16=False):
    """-> dict(max_batch, inputs=[(id, name, dims)], layers=[dict], outputs=[id]).
    round_fp16: every weight is rounded to fp16 first -- the network the reference builds from trt_weights_fp16.bin."""
    assert buf[:8] == MAGIC, "not an engine plan"
    r = _Reader(buf, 8, round_fp16)
    version, max_batch, _half2 = r.get("i"), r.get("i"), r.get("B")
    assert version == 1
    inputs = []
    for _ in range(r.get("i")):
        tid, name, _typ, dims = r.get("i"), r.str(), r.get("i"), r.dims()
        inputs.append((tid, name, dims))
    layers = []
    for _ in range(r.get("i")):
        L = {"kind": r.get("i"), "name": r.str()}
        L["in"] = [r.get("i") for _ in range(r.get("i"))]
        L["out"] = [(r.get("i"), r.str()) for _ in range(r.get("i"))]
        k = L["kind"]
        if k in (K_CONV, K_DECONV):
            L["maps"], kh, kw, sh, sw, ph, pw = r.get("i"), r.get("i"), r.get("i"), r.get("i"), r.get("i"), r.get("i"), r.get("i")
            L["ksize"], L["stride"], L["pad"] = (kh, kw), (sh, sw), (ph, pw)
            L["w"], L["b"] = r.weights(), r.weights()
        elif k == K_SCALE:
            L["mode"] = r.get("i")
            L["shift"], L["scale"], L["power"] = r.weights(), r.weights(), r.weights()
        elif k == K_ACTIVATION:
            L["act"] = r.get("i")
        elif k == K_ELTWISE:
            L["op"] = r.get("i")
        elif k == K_SHUFFLE:
            has = r.get("B")
            d = r.dims()
            L["reshape"] = d if has else None
        elif k == K_PLUGIN:
            n = r.get("q")
            r.align8()
            L["plugin"] = _parse_plugin(buf[r.p:r.p + n], round_fp16)
            r.p += n
        elif k == K_POOLING:       # the Caffe-model layers (TrailNet): pooling type, window, stride, pad, output extent
            L["pool"], L["k"], L["stride"], L["pad"], L["oh"], L["ow"] = r.get("i"), r.get("i"), r.get("i"), r.get("i"), r.get("i"), r.get("i")
        elif k == K_FC:
            L["maps"] = r.get("i")
            L["w"], L["b"] = r.weights(), r.weights()
        elif k == K_SOFTMAX:
            pass
        else:
            assert k == K_CONCAT, k
        layers.append(L)
    outputs = [r.get("i") for _ in range(r.get("i"))]
    assert r.p == len(buf), (r.p, len(buf))
    return {"max_batch": max_batch, "inputs": inputs, "layers": layers, "outputs": outputs}


def _parse_plugin(blob, round_fp16=False):
    r = _Reader(blob)
    tag = r.get("i")
    P = {"tag": tag}
    if tag == P_ELU:
        r.get("i"); r.get("B"); r.get("%di" % r.get("i"))
    elif tag == P_COSTVOL:
        r.get("i"); r.get("B")
        P["cv_type"], P["max_disp"] = r.get("i"), r.get("i")
    elif tag == P_SOFTARGMAX:
        r.get("i")
        P["is_min"] = r.get("i") == 1
    elif tag in (P_CONV3D, P_CONV3D_T):
        P["kdims"] = r.get("5i")
        P["out_dims"] = r.get("4i")
        P["stride"], P["pad_start"], P["pad_end"] = r.get("3i"), r.get("3i"), r.get("3i")
        typ, kc, bc = r.get("i"), r.get("q"), r.get("q")
        dt = np.float16 if typ == 1 else np.float32
        P["w"] = np.frombuffer(blob, dtype=dt, count=kc, offset=r.p).astype(np.float64).reshape(P["kdims"])
        r.p += kc * np.dtype(dt).itemsize
        P["b"] = np.frombuffer(blob, dtype=dt, count=bc, offset=r.p).astype(np.float64) if bc else None
        if round_fp16:
            P["w"] = _f16(P["w"])
            P["b"] = _f16(P["b"]) if P["b"] is not None else None
    elif tag == P_TRANSFORM:
        P["order"] = r.get("4i")
    elif tag == P_PADDING:
        P["planes"] = r.get("i")
    elif tag == P_SLICE:
        r.get("4i")
        P["start"], P["end"] = r.get("i"), r.get("i")
    else:
        raise ValueError("unknown plugin tag %d" % tag)
    return P


_SLAB_BYTES = 6e9      # bound on torch's im2col buffer for the float64 3-D convolutions (they are GBs per layer)


def _conv3d_slabs(x, w, b, stride, pad_start):
    """ops.conv3d over output-depth slabs (same values; the D padding is applied once, up front)."""
    n, d, c, h, wd = x.shape
    per_plane = 27.0 * c * (h // stride[1] + 1) * (wd // stride[2] + 1) * x.element_size()
    slab = max(1, int(_SLAB_BYTES / per_plane))
    pd = pad_start[0]
    d_out = (d + 2 * pd - 3) // stride[0] + 1
    if slab >= d_out:
        return ops.conv3d(x, w, b, stride, pad_start)
    xp = torch.nn.functional.pad(x, (0, 0, 0, 0, 0, 0, pd, pd))
    outs = []
    for a in range(0, d_out, slab):
        e = min(d_out, a + slab)
        outs.append(ops.conv3d(xp[:, a * stride[0]:(e - 1) * stride[0] + 3], w, b, stride, (0, pad_start[1], pad_start[2])))
    return torch.cat(outs, dim=2)


def _conv3d_transpose_slabs(y, w, b, stride, pad_start, out_dims):
    """ops.conv3d_transpose with the input depth processed in slabs (overlap-add of the slabs' full outputs)."""
    n, k, dy, hy, wy = y.shape
    c = w.shape[2]
    per_plane = 27.0 * c * hy * wy * stride[1] * stride[2] * y.element_size() * stride[0]
    slab = max(1, int(_SLAB_BYTES / per_plane))
    if slab >= dy:
        return ops.conv3d_transpose(y, w, b, stride, pad_start, out_dims)
    dx, _, hx, wx = out_dims
    s0 = stride[0]
    d_full = max((dy - 1) * s0 + 3, pad_start[0] + dx)
    acc = None
    for a in range(0, dy, slab):
        e = min(dy, a + slab)
        d_part = (e - a - 1) * s0 + 3
        part = ops.conv3d_transpose(y[:, :, a:e], w, None, stride, (0, pad_start[1], pad_start[2]), (d_part, c, hx, wx))   # [N, d_part, C, hx, wx]
        if acc is None:
            acc = part.new_zeros((n, d_full, c, hx, wx))
        acc[:, a * s0:a * s0 + d_part] += part
    out = acc[:, pad_start[0]:pad_start[0] + dx]
    if b is not None:
        out = out + b.view(1, 1, -1, 1, 1)
    return out.contiguous()


def execute(plan, feeds, dtype=torch.float64):
    """feeds: {input name: numpy [N, *dims]}  ->  {output tensor name: numpy}.  Layers run in plan (= builder) order."""
    def T(a):
        return None if a is None else torch.from_numpy(np.ascontiguousarray(a)).to(dtype)

    vals, names = {}, {}
    for tid, name, dims in plan["inputs"]:
        x = T(feeds[name])
        assert tuple(x.shape[1:]) == tuple(dims), (name, x.shape, dims)
        vals[tid], names[tid] = x, name
    with torch.no_grad():
        for L in plan["layers"]:
            x = [vals[i] for i in L["in"]]
            k = L["kind"]
            if k == K_CONV:
                cin = x[0].shape[1]
                w = T(L["w"]).view(L["maps"], cin, *L["ksize"])
                y = ops.conv2d(x[0], w, T(L["b"]), L["stride"], L["pad"])
            elif k == K_DECONV:
                cin = x[0].shape[1]
                w = T(L["w"]).view(cin, L["maps"], *L["ksize"])
                y = ops.deconv2d(x[0], w, T(L["b"]), L["stride"], L["pad"])
            elif k == K_SCALE and L["mode"] == 1:          # per channel (Caffe Scale layer): x * scale[c] + shift[c]
                assert L["power"] is None
                y = x[0]
                if L["scale"] is not None:
                    y = y * T(L["scale"]).view(1, -1, 1, 1)
                if L["shift"] is not None:
                    y = y + T(L["shift"]).view(1, -1, 1, 1)
            elif k == K_SCALE:
                assert L["mode"] == 0
                s = lambda a, dflt: float(a[0]) if a is not None else dflt
                y = ops.scale(x[0], s(L["shift"], 0.0), s(L["scale"], 1.0), s(L["power"], 1.0))
            elif k == K_ACTIVATION:
                assert L["act"] in (0, 1), "kRELU (Caffe models) / kSIGMOID (stereo builders)"
                y = torch.relu(x[0]) if L["act"] == 0 else ops.sigmoid(x[0])
            elif k == K_POOLING:
                from . import caffe
                y = torch.from_numpy(caffe._pool(x[0].numpy(), "MAX" if L["pool"] == 0 else "AVE", L["k"], L["stride"], L["pad"]))
                assert tuple(y.shape[2:]) == (L["oh"], L["ow"]), (y.shape, L["oh"], L["ow"])
            elif k == K_FC:
                x2 = x[0].reshape(x[0].shape[0], -1)
                y = x2 @ T(L["w"]).view(L["maps"], -1).t()
                if L["b"] is not None:
                    y = y + T(L["b"]).view(1, -1)
                y = y.view(y.shape[0], -1, 1, 1)
            elif k == K_SOFTMAX:
                y = torch.softmax(x[0], dim=1)
            elif k == K_ELTWISE:
                assert L["op"] == 0
                y = x[0] + x[1]
            elif k == K_CONCAT:
                y = torch.cat(x, dim=1)
            elif k == K_SHUFFLE:
                y = x[0].reshape(x[0].shape[0], *L["reshape"]) if L["reshape"] is not None else x[0]
            else:
                P = L["plugin"]
                t = P["tag"]
                if t == P_ELU:
                    y = ops.elu(x[0])
                elif t == P_COSTVOL:
                    y = ops.cost_volume(x[0], x[1], P["max_disp"]) if P["cv_type"] == 0 else ops.corr_cost_volume(x[0], x[1], P["max_disp"])
                elif t == P_SOFTARGMAX:
                    y = ops.softargmax(x[0], P["is_min"])
                elif t == P_CONV3D:
                    y = _conv3d_slabs(x[0], T(P["w"]), T(P["b"]), P["stride"], P["pad_start"])
                elif t == P_CONV3D_T:
                    y = _conv3d_transpose_slabs(x[0], T(P["w"]), T(P["b"]), P["stride"], P["pad_start"], P["out_dims"])
                elif t == P_TRANSFORM:
                    y = ops.transform(x[0], P["order"])
                elif t == P_PADDING:
                    y = ops.pad_d(x[0], P["planes"])
                else:
                    y = ops.slice_d(x[0], P["start"], P["end"])
            (oid, oname), = L["out"]
            vals[oid], names[oid] = y, oname
            for i in L["in"]:                      # free tensors nobody reads any more (the 3-D volumes are GBs in float64)
                if not any(i in M["in"] for M in plan["layers"][plan["layers"].index(L) + 1:]) and i not in plan["outputs"]:
                    vals.pop(i, None)
    return {names[o]: vals[o].numpy() for o in plan["outputs"]}
