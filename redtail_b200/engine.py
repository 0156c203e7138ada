"""Python handle on the whole-network C-ABI (include/redtail_b200_engine.h)."""
import ctypes as C

import torch

from ._lib import engine_lib
from .ops import RedtailError, RT_F32, RT_F16

MODELS = {"nvsmall": 48, "nvtiny": 24}      # half-resolution max disparity of the reference's weight sets


class StereoEngine:
    """The reference's NVSmall-family stereo net on one GPU.

    left/right: [N,3,H,W] fp32 in [0,1]  ->  disparity [N,H,W] fp32 (pixels).
    """

    def __init__(self, model, height, width, weights_path, max_batch=1, max_disp=None, weights_dtype="fp32"):
        if not torch.cuda.is_available():
            raise RedtailError("StereoEngine needs a CUDA device (there is no CPU path)")
        self.lib = engine_lib()
        self.h, self.w, self.max_batch = height, width, max_batch
        self._e = C.c_void_p()
        md = max_disp if max_disp is not None else MODELS[model]
        rc = self.lib.rt_stereo_create(model.encode(), height, width, md, str(weights_path).encode(),
                                       RT_F16 if weights_dtype == "fp16" else RT_F32, max_batch, C.byref(self._e))
        if rc != 0:
            raise RedtailError("rt_stereo_create failed (%d): %s" % (rc, self.lib.rt_stereo_last_error().decode()))

    def serialize(self):
        """-> bytes: the engine plan (ICudaEngine::serialize)."""
        n = self.lib.rt_stereo_serialize(self._e, None, 0)
        if n == 0:
            raise RedtailError("rt_stereo_serialize failed: %s" % self.lib.rt_stereo_last_error().decode())
        buf = C.create_string_buffer(n)
        self.lib.rt_stereo_serialize(self._e, buf, n)
        return buf.raw

    @classmethod
    def deserialize(cls, plan, max_batch=None):
        """Engine from a plan (IRuntime::deserializeCudaEngine + StereoDnnPluginFactory); no weight file needed.
        max_batch overrides the batch size the plan was written with."""
        if not torch.cuda.is_available():
            raise RedtailError("StereoEngine needs a CUDA device (there is no CPU path)")
        self = cls.__new__(cls)
        self.lib = engine_lib()
        self._e = C.c_void_p()
        if max_batch is None:
            rc = self.lib.rt_stereo_deserialize(plan, len(plan), C.byref(self._e))
        else:
            rc = self.lib.rt_stereo_deserialize_batch(plan, len(plan), int(max_batch), C.byref(self._e))
        if rc != 0:
            raise RedtailError("rt_stereo_deserialize failed (%d): %s" % (rc, self.lib.rt_stereo_last_error().decode()))
        self.h = self.w = self.max_batch = None      # taken from the plan by the library
        return self

    def __call__(self, left, right, out=None):
        """Device tensors in, device tensor out; asynchronous on the current torch stream."""
        assert left.is_cuda and right.is_cuda and left.dtype == torch.float32 and left.is_contiguous() and right.is_contiguous()
        n = left.shape[0]
        if out is None:
            out = torch.empty((n, left.shape[2], left.shape[3]), dtype=torch.float32, device=left.device)
        rc = self.lib.rt_stereo_enqueue(self._e, n, C.c_void_p(left.data_ptr()), C.c_void_p(right.data_ptr()),
                                        C.c_void_p(out.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream))
        if rc != 0:
            raise RedtailError("rt_stereo_enqueue failed (%d): %s" % (rc, self.lib.rt_stereo_last_error().decode()))
        return out

    def execute_host(self, left, right, out):
        """Host (ideally pinned) tensors: H2D + inference + D2H, synchronous -- the end-to-end call of the apps."""
        n = left.shape[0]
        rc = self.lib.rt_stereo_execute_host(self._e, n, C.c_void_p(left.data_ptr()), C.c_void_p(right.data_ptr()),
                                             C.c_void_p(out.data_ptr()))
        if rc != 0:
            raise RedtailError("rt_stereo_execute_host failed (%d): %s" % (rc, self.lib.rt_stereo_last_error().decode()))
        return out

    def execute_images(self, left_bgr, right_bgr, out=None, out_u16=None, u16_scale=256.0):
        """Host uint8 [N,H_src,W_src,3] BGR images (cv::imread layout) -> disparity: 8-bit H2D, GPU pre-processing
        (INTER_AREA resize to the network size, RGB, CHW, /255), inference, D2H.  out: float32 [N,H,W] host tensor and/or
        out_u16: uint16 [N,H,W] (KITTI-style PNG payload); synchronous."""
        assert left_bgr.dtype == torch.uint8 and left_bgr.dim() == 4 and left_bgr.shape[3] == 3 and left_bgr.is_contiguous()
        n, sh, sw, _ = left_bgr.shape
        rc = self.lib.rt_stereo_execute_images(self._e, n, C.c_void_p(left_bgr.data_ptr()), C.c_void_p(right_bgr.data_ptr()), sh, sw,
                                               C.c_void_p(out.data_ptr() if out is not None else 0),
                                               C.c_void_p(out_u16.data_ptr() if out_u16 is not None else 0), float(u16_scale))
        if rc != 0:
            raise RedtailError("rt_stereo_execute_images failed (%d): %s" % (rc, self.lib.rt_stereo_last_error().decode()))
        return out if out is not None else out_u16

    def profile(self, left, right):
        """-> list of (layer name, ms), measured with CUDA events around every engine step."""
        n = left.shape[0]
        out = torch.empty((n, left.shape[2], left.shape[3]), dtype=torch.float32, device=left.device)
        buf = C.create_string_buffer(1 << 16)
        torch.cuda.synchronize()
        rc = self.lib.rt_stereo_profile(self._e, n, C.c_void_p(left.data_ptr()), C.c_void_p(right.data_ptr()),
                                        C.c_void_p(out.data_ptr()), buf, len(buf))
        if rc != 0:
            raise RedtailError("rt_stereo_profile failed (%d)" % rc)
        rows = []
        for line in buf.value.decode().splitlines():
            name, ms = line.rsplit("\t", 1)
            rows.append((name, float(ms)))
        return rows

    @property
    def num_layers(self):
        return self.lib.rt_stereo_num_layers(self._e)

    def close(self):
        if getattr(self, "_e", None):
            self.lib.rt_stereo_destroy(self._e)
            self._e = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class CaffeNet:
    """A single-input Caffe model (the TrailNet S-ResNet-18 classifier) through the nvcaffeparser1-compatible parser and the
    engine: what ros/packages/caffe_ros/src/tensor_net.cpp does (loadNetwork / forward), minus the OpenCV pre-processing.

    x: [N,C,H,W] fp32 (for TrailNet: BGR, 0..255, 180x320)  ->  [N,Co,Ho,Wo] fp32 (TrailNet: [N,6,1,1] softmax outputs).
    """

    def __init__(self, prototxt, caffemodel, output_blob, input_blob="data", max_batch=1, _handle=None):
        if not torch.cuda.is_available():
            raise RedtailError("CaffeNet needs a CUDA device (there is no CPU path)")
        self.lib = engine_lib()
        self._e = C.c_void_p()
        if _handle is not None:
            self._e = _handle
        else:
            rc = self.lib.rt_caffe_create(str(prototxt).encode(), str(caffemodel).encode(), input_blob.encode(), output_blob.encode(),
                                          int(max_batch), C.byref(self._e))
            if rc != 0:
                raise RedtailError("rt_caffe_create failed (%d): %s" % (rc, self.lib.rt_stereo_last_error().decode()))
        i3, o3 = (C.c_int * 3)(), (C.c_int * 3)()
        self.lib.rt_net_dims(self._e, i3, o3)
        self.in_chw, self.out_chw = tuple(i3), tuple(o3)

    @classmethod
    def deserialize(cls, plan, max_batch=0):
        lib = engine_lib()
        h = C.c_void_p()
        rc = lib.rt_net_deserialize(plan, len(plan), int(max_batch), C.byref(h))
        if rc != 0:
            raise RedtailError("rt_net_deserialize failed (%d): %s" % (rc, lib.rt_stereo_last_error().decode()))
        return cls(None, None, None, _handle=h)

    def serialize(self):
        n = self.lib.rt_net_serialize(self._e, None, 0)
        if n == 0:
            raise RedtailError("rt_net_serialize failed")
        buf = C.create_string_buffer(n)
        self.lib.rt_net_serialize(self._e, buf, n)
        return buf.raw

    def __call__(self, x, out=None):
        assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and tuple(x.shape[1:]) == self.in_chw, (x.shape, self.in_chw)
        n = x.shape[0]
        if out is None:
            out = torch.empty((n,) + self.out_chw, dtype=torch.float32, device=x.device)
        rc = self.lib.rt_net_enqueue(self._e, n, C.c_void_p(x.data_ptr()), C.c_void_p(out.data_ptr()),
                                     C.c_void_p(torch.cuda.current_stream().cuda_stream))
        if rc != 0:
            raise RedtailError("rt_net_enqueue failed (%d): %s" % (rc, self.lib.rt_stereo_last_error().decode()))
        return out

    def execute_host(self, x, out):
        rc = self.lib.rt_net_execute_host(self._e, x.shape[0], C.c_void_p(x.data_ptr()), C.c_void_p(out.data_ptr()))
        if rc != 0:
            raise RedtailError("rt_net_execute_host failed (%d): %s" % (rc, self.lib.rt_stereo_last_error().decode()))
        return out

    def profile(self, x):
        out = torch.empty((x.shape[0],) + self.out_chw, dtype=torch.float32, device=x.device)
        buf = C.create_string_buffer(1 << 16)
        torch.cuda.synchronize()
        rc = self.lib.rt_net_profile(self._e, x.shape[0], C.c_void_p(x.data_ptr()), C.c_void_p(out.data_ptr()), buf, len(buf))
        if rc != 0:
            raise RedtailError("rt_net_profile failed (%d)" % rc)
        return [(l.rsplit("\t", 1)[0], float(l.rsplit("\t", 1)[1])) for l in buf.value.decode().splitlines()]

    @property
    def num_layers(self):
        return self.lib.rt_net_num_layers(self._e)

    def close(self):
        if getattr(self, "_e", None):
            self.lib.rt_net_destroy(self._e)
            self._e = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
