"""ctypes loader for the in-tree shared libraries; declares every symbol of include/redtail_b200.h and
include/redtail_b200_engine.h.  Fails loudly (LibraryMissing) when a library has not been built."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIBDIR = os.environ.get("REDTAIL_LIB_DIR") or os.path.join(HERE, "lib")     # REDTAIL_LIB_DIR: A/B builds (tools/)


class LibraryMissing(RuntimeError):
    pass


def lib_paths():
    return {"kernels": os.path.join(LIBDIR, "libredtail_b200.so"),
            "engine": os.path.join(LIBDIR, "libnvstereo_inference.so")}


class Conv3dDesc(C.Structure):
    _fields_ = [("transposed", C.c_int), ("k", C.c_int), ("v", C.c_int), ("c", C.c_int), ("r", C.c_int), ("s", C.c_int),
                ("stride", C.c_int * 3), ("pad", C.c_int * 3), ("in_dims", C.c_int * 4), ("out_dims", C.c_int * 4),
                ("weights_dtype", C.c_int), ("weights", C.c_void_p), ("bias", C.c_void_p), ("precision", C.c_int),
                ("fuse_elu", C.c_int), ("out_transposed", C.c_int), ("slice_d", C.c_int),
                ("in_layout", C.c_int), ("out_layout", C.c_int), ("pad_end_d", C.c_int), ("fuse_softargmax", C.c_int), ("act_params", C.c_void_p)]


class Conv2dDesc(C.Structure):
    _fields_ = [("transposed", C.c_int), ("cin", C.c_int), ("cout", C.c_int), ("r", C.c_int), ("s", C.c_int),
                ("stride", C.c_int * 2), ("pad", C.c_int * 2), ("in_h", C.c_int), ("in_w", C.c_int),
                ("weights_dtype", C.c_int), ("weights", C.c_void_p), ("bias", C.c_void_p), ("fuse_elu", C.c_int)]


class CostvolConv3dDesc(C.Structure):
    _fields_ = [("c", C.c_int), ("h", C.c_int), ("w", C.c_int), ("max_disp", C.c_int), ("k", C.c_int),
                ("weights_dtype", C.c_int), ("weights", C.c_void_p), ("bias", C.c_void_p), ("precision", C.c_int),
                ("fuse_elu", C.c_int), ("out_transposed", C.c_int), ("out_layout", C.c_int)]


_P, _I, _L, _F = C.c_void_p, C.c_int, C.c_int64, C.c_float

# name -> (restype, argtypes): the complete export list of include/redtail_b200.h
KERNEL_API = {
    "rt_version": (C.c_char_p, []),
    "rt_launch_count": (C.c_uint64, []),
    "rt_add_launch_count": (None, [C.c_uint64]),
    "rt_preprocess_bgr8": (_I, [_P, _I, _I, _I, _L, _P, _I, _I, _P]),
    "rt_disparity_to_u16": (_I, [_P, _P, _L, _F, _P]),
    "rt_write_png16": (_I, [C.c_char_p, _P, _I, _I]),
    "rt_last_kernel": (C.c_char_p, []),
    "rt_scale_channel": (_I, [_P, _P, _I, _I, _L, _P, _P, _P]),
    "rt_srelu": (_I, [_P, _P, _I, _I, _L, _P, _P, _P, _P, _P]),
    "rt_relu": (_I, [_P, _P, _L, _P]),
    "rt_pool2d": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    "rt_fully_connected": (_I, [_P, _P, _P, _P, _I, _I, _I, _P]),
    "rt_softmax_channels": (_I, [_P, _P, _I, _I, _L, _P]),
    "rt_im2col_split16": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    "rt_cost_volume": (_I, [_I, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "rt_corr_cost_volume": (_I, [_I, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "rt_elu": (_I, [_I, _P, _P, _L, _P]),
    "rt_sigmoid": (_I, [_I, _P, _P, _L, _P]),
    "rt_scale": (_I, [_I, _P, _P, _L, _F, _F, _F, _P]),
    "rt_eltwise_sum": (_I, [_I, _P, _P, _P, _L, _P]),
    "rt_convert": (_I, [_I, _P, _I, _P, _L, _P]),
    "rt_pad_planes": (_I, [_I, _P, _P, _I, _I, _L, _I, _P]),
    "rt_slice_planes": (_I, [_I, _P, _P, _I, _I, _L, _I, _I, _P]),
    "rt_transpose01": (_I, [_I, _P, _P, _I, _I, _I, _L, _P]),
    "rt_concat_channels": (_I, [_I, _P, _I, _P, _I, _P, _I, _L, _P]),
    "rt_softargmax": (_I, [_I, _I, _P, _P, _I, _I, _L, _P]),
    "rt_conv3d_create": (_I, [C.POINTER(Conv3dDesc), C.POINTER(_P)]),
    "rt_conv3d_tc_supported": (_I, [C.POINTER(Conv3dDesc)]),
    "rt_conv3d_destroy": (None, [_P]),
    "rt_cost_volume_split16": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "rt_dense_to_split16": (_I, [_P, _P, _I, _I, _I, _I, _I, _P]),
    "rt_split16_to_dense": (_I, [_P, _P, _I, _I, _I, _I, _I, _P]),
    "rt_conv3d_workspace_size": (C.c_size_t, [_P, _I]),
    "rt_conv3d_enqueue": (_I, [_P, _I, _P, _P, _P, _P, _P]),
    "rt_costvol_conv3d_supported": (_I, [C.POINTER(CostvolConv3dDesc)]),
    "rt_costvol_conv3d_create": (_I, [C.POINTER(CostvolConv3dDesc), C.POINTER(_P)]),
    "rt_costvol_conv3d_destroy": (None, [_P]),
    "rt_costvol_conv3d_workspace_size": (C.c_size_t, [_P, _I]),
    "rt_costvol_conv3d_enqueue": (_I, [_P, _I, _P, _P, _P, _P, _P]),
    "rt_conv2d_create": (_I, [C.POINTER(Conv2dDesc), C.POINTER(_P)]),
    "rt_conv2d_destroy": (None, [_P]),
    "rt_conv2d_out_dims": (None, [_P, C.POINTER(_I), C.POINTER(_I)]),
    "rt_conv2d_enqueue": (_I, [_P, _I, _P, _P, _P]),
}

# include/redtail_b200_engine.h
ENGINE_API = {
    "rt_stereo_create": (_I, [C.c_char_p, _I, _I, _I, C.c_char_p, _I, _I, C.POINTER(_P)]),
    "rt_stereo_destroy": (None, [_P]),
    "rt_stereo_enqueue": (_I, [_P, _I, _P, _P, _P, _P]),
    "rt_stereo_execute_host": (_I, [_P, _I, _P, _P, _P]),
    "rt_stereo_profile": (_I, [_P, _I, _P, _P, _P, C.c_char_p, C.c_size_t]),
    "rt_stereo_serialize": (C.c_size_t, [_P, _P, C.c_size_t]),
    "rt_stereo_deserialize": (_I, [_P, C.c_size_t, C.POINTER(_P)]),
    "rt_stereo_deserialize_batch": (_I, [_P, C.c_size_t, _I, C.POINTER(_P)]),
    "rt_stereo_execute_images": (_I, [_P, _I, _P, _P, _I, _I, _P, _P, _F]),
    "rt_stereo_num_layers": (_I, [_P]),
    "rt_stereo_device_bytes": (C.c_size_t, [_P]),
    "rt_stereo_last_error": (C.c_char_p, []),
    "rt_caffe_create": (_I, [C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, _I, C.POINTER(_P)]),
    "rt_net_destroy": (None, [_P]),
    "rt_net_dims": (_I, [_P, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "rt_net_enqueue": (_I, [_P, _I, _P, _P, _P]),
    "rt_net_execute_host": (_I, [_P, _I, _P, _P]),
    "rt_net_profile": (_I, [_P, _I, _P, _P, C.c_char_p, C.c_size_t]),
    "rt_net_serialize": (C.c_size_t, [_P, _P, C.c_size_t]),
    "rt_net_deserialize": (_I, [_P, C.c_size_t, _I, C.POINTER(_P)]),
    "rt_net_num_layers": (_I, [_P]),
    "rt_caffe_dump_plan": (C.c_size_t, [C.c_char_p, C.c_char_p, C.c_char_p, _I, _P, C.c_size_t]),
    # nvinfer1 shim factories (C linkage, include/NvInfer.h)
    "createInferBuilder_INTERNAL": (_P, [_P, _I]),
    "createInferRuntime_INTERNAL": (_P, [_P, _I]),
}

_cache = {}


def _load(kind, api):
    if kind in _cache:
        return _cache[kind]
    path = lib_paths()[kind]
    if not os.path.exists(path):
        raise LibraryMissing("%s not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                             "(or `make -C redtail_b200/csrc`)" % path)
    lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
    for name, (res, args) in api.items():
        fn = getattr(lib, name)          # AttributeError here = the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    _cache[kind] = lib
    return lib


def kernels_lib():
    return _load("kernels", KERNEL_API)


def engine_lib():
    kernels_lib()
    return _load("engine", ENGINE_API)
