"""Sparse evaluation of the final image fusion: the transposed convolutions, their concatenation and the 1x1 conv + BN + ReLU of
/root/reference/lib/net/pointnet2_msg.py:237-243, computed ONLY at the <= 4 bilinear taps of every point that Feature_Gather (:245,
:107-120) then reads -- 4 * 16384 of the 491520 pixels per scene -- instead of over the whole 64 x 384 x 1280 up-sampled canvas.
Kernels: csrc/sparse_tail.cu (tap list, counting sort by transposed-convolution phase, blend) + epnet_gemm_tf32x3_rows (row-gather
tcgen05 GEMM with a per-tile weight slice).  Eval mode only (BatchNorm folded); the dense form stays for the module path."""
import ctypes

import torch

from . import pointnet2_cuda as pc
from ._lib import LIB
from .gemm import PackedLinear

PHASES = 256  # (Y % 16, X % 16)


class SparseImageTail:
    def __init__(self, deconvs, fuse_w, fuse_b, batch, npoints, image_hw, device, align_corners=False):
        """deconvs: the ConvTranspose2d modules (kernel == stride, each kernel dividing 16), in concat order; fuse_w (Cf, sum Cout),
        fuse_b (Cf): the 1x1 fusion conv with BatchNorm AND the transposed convolutions' biases folded in (runner.py)."""
        self.H, self.W = image_hw
        self.B, self.N = batch, npoints
        self.align = int(bool(align_corners))
        self.k = [int(d.kernel_size[0]) for d in deconvs]
        self.cout = [int(d.out_channels) for d in deconvs]
        self.cin = [int(d.in_channels) for d in deconvs]
        for d, k in zip(deconvs, self.k):
            if d.kernel_size != (k, k) or d.stride != (k, k) or 16 % k or self.H % k or self.W % k:
                raise NotImplementedError("sparse tail: transposed convolutions need kernel == stride dividing 16 and the canvas")
        if len(deconvs) > 4 or any(c % 4 or c > 64 for c in self.cout):
            raise NotImplementedError("sparse tail: at most 4 levels of <= 64 output channels (multiple of 4)")
        self.hw = [(self.H // k, self.W // k) for k in self.k]
        # one weight set per phase: rows (ky, kx, o) x cin, i.e. n-tile (ky * k + kx) of BN = round16(Cout) columns
        self.level_lin = []
        for d, k in zip(deconvs, self.k):
            co, bn = d.out_channels, (d.out_channels + 15) // 16 * 16
            w = d.weight.detach().float().permute(2, 3, 1, 0)  # (ky, kx, Cout, Cin)
            wp = torch.zeros(k, k, bn, d.in_channels, device=w.device)
            wp[:, :, :co] = w
            lin = PackedLinear(wp.reshape(k * k * bn, d.in_channels), None)
            self.level_lin.append((lin, lin._pack(bn), bn))
        self.cat_width = sum(self.cout)
        self.fuse = PackedLinear(fuse_w, fuse_b)
        assert self.fuse.K == self.cat_width and self.fuse.N <= 64 and self.fuse.N % 4 == 0
        self.fuse_bn = (self.fuse.N + 15) // 16 * 16
        self.fuse_pack = self.fuse._pack(self.fuse_bn)
        slots = batch * npoints * 4
        self.slots = slots
        self.max_rows = (slots + PHASES * 127 + 127) // 128 * 128
        self.max_tiles = self.max_rows // 128
        i32 = dict(dtype=torch.int32, device=device)
        self.karr = (ctypes.c_int * 4)(*(self.k + [1] * (4 - len(self.k))))
        self.harr = (ctypes.c_int * 4)(*([h for h, _ in self.hw] + [1] * (4 - len(self.k))))
        self.warr = (ctypes.c_int * 4)(*([w for _, w in self.hw] + [1] * (4 - len(self.k))))

    def __call__(self, maps, xy_norm, out):
        """maps: the image-stream outputs (B, h_i, w_i, >= Cin_i) fp32 NHWC; xy_norm (B, N, 2) in [-1, 1]; out (B*N, ldo): columns
        [0, Cf) receive the fused image feature of every point (what the dense path's gather returns)."""
        dev = xy_norm.device
        i32 = dict(dtype=torch.int32, device=dev)
        f32 = dict(dtype=torch.float32, device=dev)
        B, N, L = self.B, self.N, len(self.k)
        tap_pix = torch.empty(self.slots, **i32)
        tap_w = torch.empty(self.slots, **f32)
        hist = torch.zeros(PHASES, **i32)
        pc._call("tail_taps", LIB.epnet_tail_taps, xy_norm, B, N, self.H, self.W, self.align, pc._f(xy_norm, "xy"), tap_pix.data_ptr(),
                 tap_w.data_ptr(), hist.data_ptr())
        start, cursor = torch.empty(PHASES, **i32), torch.empty(PHASES, **i32)
        tile_phase = torch.zeros(self.max_tiles, **i32)
        n_tiles = torch.empty(1, **i32)
        pc._call("tail_plan", LIB.epnet_tail_plan, xy_norm, hist.data_ptr(), start.data_ptr(), cursor.data_ptr(), tile_phase.data_ptr(),
                 n_tiles.data_ptr(), self.max_tiles)
        pos = torch.empty(self.slots, **i32)
        row_idx = torch.zeros((L, self.max_rows), **i32)  # padding rows of a bin read row 0 of the map: finite, never used
        pc._call("tail_scatter", LIB.epnet_tail_scatter, xy_norm, self.slots, self.H, self.W, L, self.karr, self.harr, self.warr,
                 tap_pix.data_ptr(), start.data_ptr(), cursor.data_ptr(), pos.data_ptr(), row_idx.data_ptr(), self.max_rows)
        cat = torch.empty((self.max_rows, self.cat_width), **f32)
        col = 0
        for i, (lin, pack, bn) in enumerate(self.level_lin):
            m = maps[i]
            assert m.is_contiguous() and m.shape[1:3] == self.hw[i] and m.shape[-1] >= self.cin[i]
            pc._call("gemm_tf32x3_rows", LIB.epnet_gemm_tf32x3_rows, m, self.max_rows, self.cin[i], self.cout[i], m.data_ptr(), m.shape[-1],
                     row_idx[i].data_ptr(), tile_phase.data_ptr(), self.k[i], n_tiles.data_ptr(), pack.data_ptr(), bn, None, 0,
                     cat.data_ptr() + 4 * col, self.cat_width)
            col += self.cout[i]
        fused = torch.empty((self.max_rows, self.fuse.N), **f32)
        pc._call("gemm_tf32x3_rows", LIB.epnet_gemm_tf32x3_rows, cat, self.max_rows, self.cat_width, self.fuse.N, cat.data_ptr(), self.cat_width,
                 None, None, 1, n_tiles.data_ptr(), self.fuse_pack.data_ptr(), self.fuse_bn, self.fuse.bias.data_ptr(), 1,
                 fused.data_ptr(), self.fuse.N)
        pc._call("tail_blend", LIB.epnet_tail_blend, fused, B * N, self.fuse.N, fused.data_ptr(), self.fuse.N, pos.data_ptr(), tap_w.data_ptr(),
                 out.data_ptr(), out.stride(0))
        return out
