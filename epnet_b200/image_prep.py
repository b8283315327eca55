"""Image input preparation on the device -- replaces the reference's host-side float64 normalisation + zero padding
(/root/reference/lib/datasets/kitti_dataset.py:37-57) and its float64 upload / .float() / permute
(/root/reference/lib/net/train_functions.py:37) with ONE kernel over the decoded uint8 image (csrc/image_prep.cu).
Values are bit-identical to the reference's: the kernel evaluates (v / 255.0 - mean) / std in float64 and rounds once."""
import ctypes

import torch

from . import pointnet2_cuda as pc
from ._lib import LIB

MEAN = (0.485, 0.456, 0.406)  # kitti_dataset.py:24
STD = (0.229, 0.224, 0.225)   # kitti_dataset.py:25
CANVAS_HW = (384, 1280)       # kitti_dataset.py:54


def _dbl3(v):
    return (ctypes.c_double * 3)(*[float(x) for x in v])


def normalise_pad(img_u8, sizes=None, out_hw=CANVAS_HW, nhwc4=None, nchw=None, mean=MEAN, std=STD):
    """img_u8 (B,h,w,3) uint8 RGB on the device (rows may be strided; pixels interleaved) -> the zero-padded, normalised canvas.
    sizes (B,2) int32 device tensor {rows, cols} decoded per scene, or None = all h x w.
    Fills `nhwc4` (B,H,W,4) fp32 and/or `nchw` (B,3,H,W) fp32; with neither given, allocates and returns the (B,3,H,W) tensor the
    reference's model takes as `img`."""
    if img_u8.dtype != torch.uint8 or img_u8.dim() != 4 or img_u8.shape[-1] != 3 or not img_u8.is_cuda:
        raise ValueError("img_u8 must be a CUDA uint8 tensor (B,h,w,3)")
    if img_u8.stride(-1) != 1 or img_u8.stride(-2) != 3:
        raise ValueError("pixels must be interleaved RGB bytes")
    b, h, w, _ = img_u8.shape
    H, W = out_hw
    if h > H or w > W:
        raise ValueError("image %dx%d exceeds the %dx%d canvas" % (h, w, H, W))
    if nhwc4 is None and nchw is None:
        nchw = torch.empty((b, 3, H, W), dtype=torch.float32, device=img_u8.device)
    if nhwc4 is not None and (nhwc4.shape != (b, H, W, 4) or not nhwc4.is_contiguous()):
        raise ValueError("nhwc4 must be contiguous (B,H,W,4)")
    if nchw is not None and (nchw.shape != (b, 3, H, W) or not nchw.is_contiguous()):
        raise ValueError("nchw must be contiguous (B,3,H,W)")
    pc._call("image_prep_u8", LIB.epnet_image_prep_u8, img_u8, b, h, w, img_u8.stride(1), img_u8.stride(0), img_u8.data_ptr(),
             None if sizes is None else pc._i(sizes, "sizes"), H, W, _dbl3(mean), _dbl3(std),
             None if nhwc4 is None else pc._f(nhwc4, "nhwc4"), None if nchw is None else pc._f(nchw, "nchw"))
    return nchw if nchw is not None else nhwc4


def nchw_to_nhwc4(image, out=None):
    """image (B,3,H,W) fp32 contiguous -> (B,H,W,4) fp32 with channel 3 = 0"""
    b, c, H, W = image.shape
    if c != 3:
        raise ValueError("image must have 3 channels")
    if out is None:
        out = torch.empty((b, H, W, 4), dtype=torch.float32, device=image.device)
    pc._call("image_nchw_to_nhwc4", LIB.epnet_image_nchw_to_nhwc4, image, b, H, W, pc._f(image, "image"), pc._f(out, "out"))
    return out
