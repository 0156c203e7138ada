"""Op-level CPU restatement of the reference plugin library (TEST INFRASTRUCTURE, see __init__).

All tensors are torch CPU tensors with a leading batch dim N (samples are independent; the
reference's plugin path is effectively batch 1 -- SURVEY.md D5).  Layout names follow the
reference (`stereoDNN/scripts/data_converters.py:13-58`):

  2-D activation          [N, C, H, W]
  3-D activation "NDCHW"  [N, D, C, H, W]      between plugins
  3-D conv output "NCDHW" [N, K, D, H, W]      Conv3DPlugin output / Conv3DTransposePlugin input
  3-D weights  "KVCRS"    [K, V, C, R, S]      (`data_converters.py:49-58`)
  2-D weights  "KCRS"     [K, C, R, S]; 2-D deconv weights [Cin, Cout, R, S]
"""
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------
# Cost volume.  Reference: lib/kernels.cu:50-97 (K1 copy + K2 shifted copy), spec
# scripts/test_data_generator.py:223-240.  out[d, c] = L[c];  out[d, C+c, h, w] = R[c, h, w-d] or 0.
# --------------------------------------------------------------------------------------
def cost_volume(left, right, max_disp):
    n, c, h, w = left.shape
    out = left.new_zeros((n, max_disp, 2 * c, h, w))
    for d in range(max_disp):
        out[:, d, :c] = left
        if d < w:
            out[:, d, c:, :, d:] = right[:, :, :, : w - d]
    return out


# Correlation cost volume.  Reference: lib/kernels.cu:168-200 (K4), spec
# scripts/test_data_generator.py:242-259.  out[d,h,w] = sum_c L[c,h,w] * R[c,h,w-d] (0 for w<d).
def corr_cost_volume(left, right, max_disp):
    n, c, h, w = left.shape
    out = left.new_zeros((n, max_disp, h, w))
    for d in range(max_disp):
        if d < w:
            out[:, d, :, d:] = (left[:, :, :, d:] * right[:, :, :, : w - d]).sum(dim=1)
    return out


# --------------------------------------------------------------------------------------
# Conv3DPlugin, kTensorFlow flavour.  Reference: lib/conv3d_plugin.cpp:74-100,187-216 and the
# reshape trick lib/conv_utils.cpp:14-81.  cuDNN cross-correlation with SYMMETRIC padding
# = pad_start on every axis (pad_end is only asserted, conv3d_plugin.cpp:48-52; an extra
# D-end plane, when TF-SAME is asymmetric, is supplied by PaddingPlugin before the conv).
#   x [N, D, C, H, W], w [K, V, C, R, S], b [K] | None  ->  y [N, K, Do, Ho, Wo]
# --------------------------------------------------------------------------------------
def conv3d(x, w, b, stride, pad_start):
    x5 = x.permute(0, 2, 1, 3, 4)            # N C D H W
    w5 = w.permute(0, 2, 1, 3, 4)            # K C V R S
    return F.conv3d(x5, w5, b, stride=tuple(stride), padding=tuple(pad_start))


# --------------------------------------------------------------------------------------
# Conv3DTransposePlugin.  Reference: lib/conv3d_transpose_plugin.cpp:86-114,205-243
# (cudnnConvolutionBackwardData of the conv above + bias over C via kernels.cu:292-308).
#   y [N, K, Dy, Hy, Wy], w [K, V, C, R, S], b [C] | None, out_dims (Dx, C, Hx, Wx)
#   x[dx, c, hx, wx] = b[c] + sum_{k,v,r,s : dx+pd = dy*sd+v, ...} w[k,v,c,r,s] * y[k,dy,hy,wy]
#   -> x [N, Dx, C, Hx, Wx]   (the generator may inflate Dx by one; SlicePlugin drops it,
#      scripts/tensorrt_model_builder.py:422-434)
# --------------------------------------------------------------------------------------
def conv3d_transpose(y, w, b, stride, pad_start, out_dims):
    dx, c, hx, wx = out_dims
    w5 = w.permute(0, 2, 1, 3, 4)            # (in=K, out=C, V, R, S)
    full = F.conv_transpose3d(y, w5, None, stride=tuple(stride))     # N C Dfull Hfull Wfull
    need = [pad_start[0] + dx, pad_start[1] + hx, pad_start[2] + wx]
    ext = [max(0, need[i] - full.shape[2 + i]) for i in range(3)]
    if any(ext):
        full = F.pad(full, (0, ext[2], 0, ext[1], 0, ext[0]))
    out = full[:, :, pad_start[0]:need[0], pad_start[1]:need[1], pad_start[2]:need[2]]
    if b is not None:
        out = out + b.view(1, -1, 1, 1, 1)
    return out.permute(0, 2, 1, 3, 4).contiguous()


# ELU, alpha = 1.  Reference: lib/elu_plugin.cpp:87-99,123-135 (cudnnActivationForward ELU).
def elu(x):
    return torch.where(x > 0, x, torch.expm1(x))


# Soft-arg{max,min} over D.  Reference: lib/softargmax_plugin.cpp:167-205 (copy, *-1, softmax
# ACCURATE over channel, * index, reduce-sum); spec scripts/test_data_generator.py:296-312.
#   x [N, D, 1, H, W] or [N, D, H, W]  ->  [N, 1, H, W]
def softargmax(x, is_min):
    if x.dim() == 5:
        assert x.shape[2] == 1
        x = x[:, :, 0]
    z = -x if is_min else x
    p = torch.softmax(z, dim=1)
    idx = torch.arange(x.shape[1], dtype=x.dtype).view(1, -1, 1, 1)
    return (p * idx).sum(dim=1, keepdim=True)


# PaddingPlugin: zero planes appended to the outermost dim of a 4-D tensor.
# Reference: lib/padding_plugin.cpp:18-31,79-94.
def pad_d(x, pad_end):
    return F.pad(x, (0, 0, 0, 0, 0, 0, 0, pad_end))


# SlicePlugin: [start, end) on the outermost dim.  Reference: lib/slice_plugin.cpp:20-36,80-92.
def slice_d(x, start, end):
    return x[:, start:end].contiguous()


# TransformPlugin: 4-D permutation; only {1,0,2,3} is used.  Reference: lib/transform_plugin.cpp:94-108.
def transform(x, order=(1, 0, 2, 3)):
    return x.permute(0, *[o + 1 for o in order]).contiguous()


# ---------------------------------------------------------------------------------------
# TRT-native layers the generated builders call.
# ---------------------------------------------------------------------------------------
# IScaleLayer kUNIFORM: (x * scale + shift) ^ power   (sample_app/nvsmall_1025x321_net.cpp:36-45).
def scale(x, shift, scl, power):
    y = x * scl + shift
    return y if power == 1.0 else y ** power


# IConvolutionLayer: KCRS weights, symmetric pad (nvsmall_1025x321_net.cpp:48-53).
def conv2d(x, w, b, stride, pad):
    return F.conv2d(x, w, b, stride=tuple(stride), padding=tuple(pad))


# IDeconvolutionLayer: weights [Cin, Cout, R, S], out = (in-1)*s + k - 2p
# (resnet18_2D_513x257_net.cpp:613; scripts/tensorrt_model_builder.py:263,275-278).
def deconv2d(x, w, b, stride, pad):
    return F.conv_transpose2d(x, w, b, stride=tuple(stride), padding=tuple(pad))


def sigmoid(x):
    return torch.sigmoid(x)
