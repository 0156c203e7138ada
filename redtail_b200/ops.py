"""Python mirror of the plugin-level C-ABI (include/redtail_b200.h) on torch CUDA tensors.

Argument meaning follows the reference plugins (stereoDNN/lib/*_plugin.cpp); tensors carry a leading batch dim.
Every function launches the library's CUDA kernels on the current torch stream; nothing here computes on the host.
"""
import ctypes as C

import numpy as np
import torch

from ._lib import kernels_lib, Conv3dDesc, Conv2dDesc, CostvolConv3dDesc

RT_F32, RT_F16 = 0, 1
PREC_FP32, PREC_FP16, PREC_SIMT = 0, 1, 2
LAYOUT_DENSE, LAYOUT_SPLIT16 = 0, 1


class RedtailError(RuntimeError):
    pass


def _check(rc, what):
    if rc != 0:
        raise RedtailError("%s failed with status %d" % (what, rc))


def _dt(t):
    if t.dtype == torch.float32:
        return RT_F32
    if t.dtype == torch.float16:
        return RT_F16
    raise TypeError("unsupported dtype %s" % t.dtype)


def _dev(*ts):
    for t in ts:
        if t is not None and not (t.is_cuda and t.is_contiguous()):
            raise ValueError("redtail_b200 ops need contiguous CUDA tensors (there is no CPU path)")


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def cost_volume(left, right, max_disp):
    """[N,C,H,W] x2 -> [N,D,2C,H,W]   (CostVolumePlugin kDefault)."""
    _dev(left, right)
    n, c, h, w = left.shape
    out = torch.empty((n, max_disp, 2 * c, h, w), dtype=left.dtype, device=left.device)
    _check(kernels_lib().rt_cost_volume(_dt(left), _p(left), _p(right), _p(out), n, c, h, w, max_disp, _stream()), "rt_cost_volume")
    return out


def cost_volume_split16(left, right, max_disp):
    """[N,C,H,W] fp32 x2 -> RT_LAYOUT_SPLIT16 cost volume: half tensor [N,2(hi|lo),D,H,W,2C]."""
    _dev(left, right)
    n, c, h, w = left.shape
    out = torch.empty((n, 2, max_disp, h, w, 2 * c), dtype=torch.float16, device=left.device)
    _check(kernels_lib().rt_cost_volume_split16(_p(left), _p(right), _p(out), n, c, h, w, max_disp, _stream()), "rt_cost_volume_split16")
    return out


def dense_to_split16(x):
    """dense fp32 [N,D,C,H,W] -> half [N,2(hi|lo),D,H,W,C] with x = hi + lo/2048."""
    _dev(x)
    n, d, c, h, w = x.shape
    out = torch.empty((n, 2, d, h, w, c), dtype=torch.float16, device=x.device)
    _check(kernels_lib().rt_dense_to_split16(_p(x), _p(out), n, d, c, h, w, _stream()), "rt_dense_to_split16")
    return out


def split16_to_dense(x):
    """half [N,2,D,H,W,C] -> dense fp32 [N,D,C,H,W]."""
    _dev(x)
    n, _, d, h, w, c = x.shape
    out = torch.empty((n, d, c, h, w), dtype=torch.float32, device=x.device)
    _check(kernels_lib().rt_split16_to_dense(_p(x), _p(out), n, d, c, h, w, _stream()), "rt_split16_to_dense")
    return out


def corr_cost_volume(left, right, max_disp):
    """[N,C,H,W] x2 -> [N,D,H,W]   (CostVolumePlugin kCorrelation)."""
    _dev(left, right)
    n, c, h, w = left.shape
    out = torch.empty((n, max_disp, h, w), dtype=left.dtype, device=left.device)
    _check(kernels_lib().rt_corr_cost_volume(_dt(left), _p(left), _p(right), _p(out), n, c, h, w, max_disp, _stream()), "rt_corr_cost_volume")
    return out


def elu(x):
    _dev(x)
    y = torch.empty_like(x)
    _check(kernels_lib().rt_elu(_dt(x), _p(x), _p(y), x.numel(), _stream()), "rt_elu")
    return y


def sigmoid(x):
    _dev(x)
    y = torch.empty_like(x)
    _check(kernels_lib().rt_sigmoid(_dt(x), _p(x), _p(y), x.numel(), _stream()), "rt_sigmoid")
    return y


def scale(x, shift, scl, power):
    _dev(x)
    y = torch.empty_like(x)
    _check(kernels_lib().rt_scale(_dt(x), _p(x), _p(y), x.numel(), shift, scl, power, _stream()), "rt_scale")
    return y


def eltwise_sum(a, b):
    _dev(a, b)
    y = torch.empty_like(a)
    _check(kernels_lib().rt_eltwise_sum(_dt(a), _p(a), _p(b), _p(y), a.numel(), _stream()), "rt_eltwise_sum")
    return y


def convert(x, dtype):
    _dev(x)
    y = torch.empty(x.shape, dtype=dtype, device=x.device)
    _check(kernels_lib().rt_convert(_dt(x), _p(x), _dt(y), _p(y), x.numel(), _stream()), "rt_convert")
    return y


def pad_d(x, pad_end):
    """[N,D,...] -> [N,D+pad_end,...], zero planes appended (PaddingPlugin)."""
    _dev(x)
    n, d = x.shape[:2]
    plane = int(np.prod(x.shape[2:]))
    y = torch.empty((n, d + pad_end) + tuple(x.shape[2:]), dtype=x.dtype, device=x.device)
    _check(kernels_lib().rt_pad_planes(_dt(x), _p(x), _p(y), n, d, plane, pad_end, _stream()), "rt_pad_planes")
    return y


def slice_d(x, start, end):
    """[N,D,...] -> [N,end-start,...] (SlicePlugin)."""
    _dev(x)
    n, d = x.shape[:2]
    plane = int(np.prod(x.shape[2:]))
    y = torch.empty((n, end - start) + tuple(x.shape[2:]), dtype=x.dtype, device=x.device)
    _check(kernels_lib().rt_slice_planes(_dt(x), _p(x), _p(y), n, d, plane, start, end, _stream()), "rt_slice_planes")
    return y


def transform(x):
    """[N,A,B,H,W] -> [N,B,A,H,W] (TransformPlugin {1,0,2,3})."""
    _dev(x)
    n, a, b = x.shape[:3]
    inner = int(np.prod(x.shape[3:]))
    y = torch.empty((n, b, a) + tuple(x.shape[3:]), dtype=x.dtype, device=x.device)
    _check(kernels_lib().rt_transpose01(_dt(x), _p(x), _p(y), n, a, b, inner, _stream()), "rt_transpose01")
    return y


def concat_channels(a, b):
    _dev(a, b)
    n, ca = a.shape[:2]
    cb = b.shape[1]
    inner = int(np.prod(a.shape[2:]))
    y = torch.empty((n, ca + cb) + tuple(a.shape[2:]), dtype=a.dtype, device=a.device)
    _check(kernels_lib().rt_concat_channels(_dt(a), _p(a), ca, _p(b), cb, _p(y), n, inner, _stream()), "rt_concat_channels")
    return y


def softargmax(x, is_min):
    """[N,D,1,H,W] or [N,D,H,W] -> [N,1,H,W] (SoftargmaxPlugin)."""
    _dev(x)
    if x.dim() == 5:
        assert x.shape[2] == 1
        x = x[:, :, 0]
    n, d, h, w = x.shape
    y = torch.empty((n, 1, h, w), dtype=x.dtype, device=x.device)
    _check(kernels_lib().rt_softargmax(_dt(x), int(bool(is_min)), _p(x), _p(y), n, d, h * w, _stream()), "rt_softargmax")
    return y


class Conv3d:
    """Conv3DPlugin / Conv3DTransposePlugin plan (weights are repacked and uploaded once, like the plugin's configure()).

    conv      : x [N,D,C,H,W]  -> y [N,K,Do,Ho,Wo]   (or [N,Do,K,Ho,Wo] with out_transposed)
    transposed: y [N,K,Dy,Hy,Wy] -> x [N,Dx-slice_d,C,Hx,Wx]   with out_dims = (Dx,C,Hx,Wx)
    """

    def __init__(self, weights, bias, stride, pad_start, in_dims, out_dims=None, transposed=False,
                 precision=PREC_FP32, fuse_elu=False, out_transposed=False, slice_d=0,
                 in_layout=LAYOUT_DENSE, out_layout=LAYOUT_DENSE, pad_end_d=0, fuse_softargmax=0, act=None):
        w = np.ascontiguousarray(weights)
        assert w.ndim == 5 and w.dtype in (np.float32, np.float16)
        b = None if bias is None else np.ascontiguousarray(bias).astype(w.dtype)
        k, v, c, r, s = w.shape
        d = Conv3dDesc()
        d.transposed = int(transposed)
        d.k, d.v, d.c, d.r, d.s = k, v, c, r, s
        d.stride[:] = list(stride)
        d.pad[:] = list(pad_start)
        d.in_dims[:] = list(in_dims)
        if not transposed:
            sp = (in_dims[0] + pad_end_d, in_dims[2], in_dims[3])
            kk = (v, r, s)
            o = [(sp[i] + 2 * pad_start[i] - kk[i]) // stride[i] + 1 for i in range(3)]
            out_dims = (k, o[0], o[1], o[2])
        d.out_dims[:] = list(out_dims)
        d.weights_dtype = RT_F32 if w.dtype == np.float32 else RT_F16
        d.weights = w.ctypes.data
        d.bias = b.ctypes.data if b is not None else None
        d.precision = precision
        d.fuse_elu = int(fuse_elu)
        d.out_transposed = int(out_transposed)
        d.slice_d = int(slice_d)
        d.in_layout, d.out_layout, d.pad_end_d = int(in_layout), int(out_layout), int(pad_end_d)
        d.fuse_softargmax = int(fuse_softargmax)     # transposed, one output channel: 1 soft-argmin / 2 soft-argmax -> y [N,Hx,Wx]
        self._act = None if act is None else np.ascontiguousarray(act, dtype=np.float32)    # [4,K]: s1, b1, s2, b2 (fused S-ReLU)
        d.act_params = self._act.ctypes.data if self._act is not None else None
        self.desc = d
        self.transposed = transposed
        self.out_dims = tuple(out_dims)
        self._plan = C.c_void_p()
        rc = kernels_lib().rt_conv3d_create(C.byref(d), C.byref(self._plan))
        if rc != 0:
            raise RedtailError("rt_conv3d_create failed with status %d" % rc)
        self._ws = None

    def __call__(self, x, skip=None):
        """Dense layouts: fp32 tensors as documented above.  RT_LAYOUT_SPLIT16: half tensors [N,2,D,H,W,C]."""
        _dev(x, skip)
        n = x.shape[0]
        if self.desc.in_layout == LAYOUT_DENSE:
            assert x.dtype == torch.float32
            assert tuple(x.shape[1:]) == tuple(self.desc.in_dims), (x.shape, tuple(self.desc.in_dims))
        else:
            assert x.dtype == torch.float16 and x.shape[1] == 2
        od = self.out_dims
        if self.desc.out_layout == LAYOUT_SPLIT16:
            if self.transposed:
                shape = (n, 2, od[0] - self.desc.slice_d, od[2], od[3], od[1])
            else:
                shape = (n, 2, od[1], od[2], od[3], od[0])
            y = torch.empty(shape, dtype=torch.float16, device=x.device)
        else:
            if self.desc.fuse_softargmax:
                shape = (n, od[2], od[3])
            elif self.transposed:
                shape = (n, od[0] - self.desc.slice_d, od[1], od[2], od[3])
            elif self.desc.out_transposed:
                shape = (n, od[1], od[0], od[2], od[3])
            else:
                shape = (n,) + od
            y = torch.empty(shape, dtype=torch.float32, device=x.device)
        lib = kernels_lib()
        need = lib.rt_conv3d_workspace_size(self._plan, n)
        if need and (self._ws is None or self._ws.numel() < need):
            self._ws = torch.empty(need, dtype=torch.uint8, device=x.device)
        _check(lib.rt_conv3d_enqueue(self._plan, n, _p(x), _p(skip), _p(y), _p(self._ws) if need else None, _stream()),
               "rt_conv3d_enqueue")
        return y

    def __del__(self):
        try:
            if getattr(self, "_plan", None):
                kernels_lib().rt_conv3d_destroy(self._plan)
        except Exception:
            pass


class CostVolumeConv3d:
    """Fused CostVolumePlugin(kDefault) -> Conv3DPlugin(3x3x3, stride 1, pad 1) [-> Transform] [-> ELU].

    left, right [N,C,H,W] fp32 -> [N,K,D,H,W] ([N,D,K,H,W] with out_transposed) fp32, or half [N,2,D,H,W,K] (split16).
    weights KVCRS [K,3,2C,3,3].  The cost volume is never materialised (include/redtail_b200.h, rt_costvol_conv3d_*).
    """

    def __init__(self, weights, bias, in_chw, max_disp, precision=PREC_FP32, fuse_elu=False, out_transposed=False,
                 out_layout=LAYOUT_DENSE):
        w = np.ascontiguousarray(weights)
        assert w.ndim == 5 and w.dtype in (np.float32, np.float16)
        b = None if bias is None else np.ascontiguousarray(bias).astype(w.dtype)
        c, h, wd = in_chw
        assert w.shape[1:] == (3, 2 * c, 3, 3), w.shape
        d = CostvolConv3dDesc()
        d.c, d.h, d.w, d.max_disp, d.k = c, h, wd, int(max_disp), w.shape[0]
        d.weights_dtype = RT_F32 if w.dtype == np.float32 else RT_F16
        d.weights = w.ctypes.data
        d.bias = b.ctypes.data if b is not None else None
        d.precision, d.fuse_elu, d.out_transposed, d.out_layout = precision, int(fuse_elu), int(out_transposed), int(out_layout)
        self.desc = d
        self._plan = C.c_void_p()
        rc = kernels_lib().rt_costvol_conv3d_create(C.byref(d), C.byref(self._plan))
        if rc != 0:
            raise RedtailError("rt_costvol_conv3d_create failed with status %d" % rc)
        self._ws = None

    def __call__(self, left, right):
        _dev(left, right)
        d = self.desc
        n = left.shape[0]
        assert left.dtype == torch.float32 and tuple(left.shape[1:]) == (d.c, d.h, d.w) and right.shape == left.shape
        if d.out_layout == LAYOUT_SPLIT16:
            y = torch.empty((n, 2, d.max_disp, d.h, d.w, d.k), dtype=torch.float16, device=left.device)
        elif d.out_transposed:
            y = torch.empty((n, d.max_disp, d.k, d.h, d.w), dtype=torch.float32, device=left.device)
        else:
            y = torch.empty((n, d.k, d.max_disp, d.h, d.w), dtype=torch.float32, device=left.device)
        lib = kernels_lib()
        need = lib.rt_costvol_conv3d_workspace_size(self._plan, n)
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=left.device)
        _check(lib.rt_costvol_conv3d_enqueue(self._plan, n, _p(left), _p(right), _p(y), _p(self._ws), _stream()),
               "rt_costvol_conv3d_enqueue")
        return y

    def __del__(self):
        try:
            if getattr(self, "_plan", None):
                kernels_lib().rt_costvol_conv3d_destroy(self._plan)
        except Exception:
            pass


class Conv2d:
    """IConvolutionLayer / IDeconvolutionLayer plan: x [N,Cin,H,W] -> y [N,Cout,Ho,Wo]."""

    def __init__(self, weights, bias, stride, pad, in_hw, transposed=False, fuse_elu=False):
        w = np.ascontiguousarray(weights, dtype=np.float32)
        b = None if bias is None else np.ascontiguousarray(bias, dtype=np.float32)
        d = Conv2dDesc()
        d.transposed = int(transposed)
        if transposed:
            d.cin, d.cout = w.shape[0], w.shape[1]
        else:
            d.cout, d.cin = w.shape[0], w.shape[1]
        d.r, d.s = w.shape[2], w.shape[3]
        d.stride[:] = list(stride)
        d.pad[:] = list(pad)
        d.in_h, d.in_w = in_hw
        d.weights_dtype = RT_F32
        d.weights = w.ctypes.data
        d.bias = b.ctypes.data if b is not None else None
        d.fuse_elu = int(fuse_elu)
        self.desc = d
        self._plan = C.c_void_p()
        rc = kernels_lib().rt_conv2d_create(C.byref(d), C.byref(self._plan))
        if rc != 0:
            raise RedtailError("rt_conv2d_create failed with status %d" % rc)
        oh, ow = C.c_int(), C.c_int()
        kernels_lib().rt_conv2d_out_dims(self._plan, C.byref(oh), C.byref(ow))
        self.out_hw = (oh.value, ow.value)

    def __call__(self, x):
        _dev(x)
        n = x.shape[0]
        y = torch.empty((n, self.desc.cout) + self.out_hw, dtype=torch.float32, device=x.device)
        _check(kernels_lib().rt_conv2d_enqueue(self._plan, n, _p(x), _p(y), _stream()), "rt_conv2d_enqueue")
        return y

    def __del__(self):
        try:
            if getattr(self, "_plan", None):
                kernels_lib().rt_conv2d_destroy(self._plan)
        except Exception:
            pass


def launch_count():
    return int(kernels_lib().rt_launch_count())


def last_kernel():
    return kernels_lib().rt_last_kernel().decode()


def preprocess_bgr8(images, out_h, out_w):
    """uint8 [N,H,W,3] BGR (what cv::imread returns) on the device -> float32 [N,3,out_h,out_w] RGB in [0,1]:
    readImgFile of sample_app/main.cpp:83-98 (float, INTER_AREA resize, BGR->RGB, CHW, /255) as one kernel."""
    _dev(images)
    assert images.dtype == torch.uint8 and images.dim() == 4 and images.shape[3] == 3 and images.is_contiguous()
    n, h, w, _ = images.shape
    out = torch.empty((n, 3, out_h, out_w), dtype=torch.float32, device=images.device)
    _check(kernels_lib().rt_preprocess_bgr8(_p(images), n, h, w, 3 * w, _p(out), out_h, out_w, _stream()), "rt_preprocess_bgr8")
    return out


def disparity_to_u16(disp, scale=256.0):
    """float32 disparity -> uint16 payload of the KITTI-style PNG (sample_app/main.cpp:317-330)."""
    _dev(disp)
    assert disp.dtype == torch.float32 and disp.is_contiguous()
    out = torch.empty(disp.shape, dtype=torch.uint16, device=disp.device)
    _check(kernels_lib().rt_disparity_to_u16(_p(disp), _p(out), disp.numel(), float(scale), _stream()), "rt_disparity_to_u16")
    return out


def write_png16(path, pixels):
    """HOST: numpy uint16 [H,W] -> 16-bit greyscale PNG."""
    import numpy as np
    a = np.ascontiguousarray(pixels, dtype=np.uint16)
    assert a.ndim == 2
    _check(kernels_lib().rt_write_png16(str(path).encode(), a.ctypes.data_as(C.c_void_p), a.shape[0], a.shape[1]), "rt_write_png16")
