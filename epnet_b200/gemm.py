"""Host side of the tcgen05 3xTF32 GEMM (csrc/gemm_tf32x3.cu): weight packing and the launch wrapper."""
import contextlib
import os as _os

import torch

from . import pointnet2_cuda as pc
from ._lib import LIB

BK = 32

# "latency": a launch whose grid would leave most SMs idle is cut into narrower column tiles (shorter critical path of a single
# forward); "throughput": keep the widest tiles (least shared-memory traffic per flop), right when several batches are in flight
# and the GPU is full anyway.  Measured on the backbone: latency tiles 3.51 -> 3.25 ms per single batch, but -1.3 % scenes/s
# with six batches in flight.
_TILE_POLICY = "latency"

# Wide tiles (BN > 64) on the FP16 two-term split (csrc/gemm_tf32x3.cu, gemm_f16x3_kernel) instead of the TF32 split: same 22
# significand bits, twice the MMA rate, half the operand bytes.  Needs |x|, |w| < 65504 (fp16 range).  Guards:
#   * weights: a layer whose (folded) weights reach F16_LIMIT never takes the FP16 kernel (PackedLinear.f16_ok, checked at pack time);
#   * activations: every GEMM epilogue raises a device flag when it writes |y| > 6e4 or a non-finite value (OverflowFlag below);
#     the runner reads it with the result and re-runs on the TF32 split (range = fp32's) -- runner.py.
F16_WIDE = _os.environ.get("EPNET_F16_WIDE", "1") != "0"
F16_LIMIT = 6.0e4
BK16 = 64


@contextlib.contextmanager
def f16_split(enabled):
    """enable / disable the FP16-split kernels for the launches (and CUDA-graph captures) made inside the block"""
    global F16_WIDE
    prev, F16_WIDE = F16_WIDE, bool(enabled) and F16_WIDE
    try:
        yield
    finally:
        F16_WIDE = prev


class OverflowFlag:
    """Host view of the library's per-device range-guard flag (epnet_gemm_overflow_read / _reset)."""

    def __init__(self, device):
        self.device = torch.device(device)
        self.host = torch.zeros(1, dtype=torch.int32).pin_memory()

    def read_async(self):
        """enqueue flag -> pinned host word on the current stream of the device"""
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device)
            from ._lib import check
            check(LIB.epnet_gemm_overflow_read(self.host.data_ptr(), stream.cuda_stream), "gemm_overflow_read")

    def reset(self):
        with torch.cuda.device(self.device):
            from ._lib import check
            check(LIB.epnet_gemm_overflow_reset(torch.cuda.current_stream(self.device).cuda_stream), "gemm_overflow_reset")
        self.host.zero_()

    def value(self):
        """the last value read (call after synchronising the stream read_async was issued on)"""
        return int(self.host.item())


@contextlib.contextmanager
def tile_policy(policy):
    global _TILE_POLICY
    assert policy in ("latency", "throughput")
    prev, _TILE_POLICY = _TILE_POLICY, policy
    try:
        yield
    finally:
        _TILE_POLICY = prev


def choose_bn(n, rows=None):
    """Output-channel tile: equal tiles of <= 64 columns for N <= 128, else of <= 256 columns (rounded up to 16).  With the row count known,
    a grid that would leave most of the 148 SMs idle (few 128-row tiles) is cut into narrower column tiles, down to 64."""
    # up to 128 columns: tiles of <= 64, which run on the A-from-TMEM kernel (two of them re-read X, but measured faster than one
    # 128-column tile of the shared-memory-operand kernel: 835 -> 849 scenes/s on the backbone)
    tiles = (n + 63) // 64 if n <= 128 else (n + 255) // 256
    per = (n + tiles - 1) // tiles
    bn = ((per + 15) // 16) * 16
    if rows is not None and _TILE_POLICY == "latency":
        m_tiles = (rows + 127) // 128
        for cand in (128, 64):
            if cand < bn and m_tiles * ((n + bn - 1) // bn) < 100:
                bn = cand
    return bn


class Planes:
    """An activation stored as the two FP16 planes of the FP16 split (x = h1 + 2^-11 h2), written by the producing layer's epilogue and
    read by the consuming layer with TMA (csrc/gemm_tf32x3.cu, gemm_f16x3_tma_kernel).  h1, h2: (..., C) float16, last stride 1."""

    def __init__(self, h1, h2):
        assert h1.dtype == torch.float16 and h2.dtype == torch.float16 and h1.shape == h2.shape and h1.stride() == h2.stride()
        self.h1, self.h2 = h1, h2

    @classmethod
    def empty(cls, shape, device):
        buf = torch.empty((2,) + tuple(shape), dtype=torch.float16, device=device)
        return cls(buf[0], buf[1])

    @property
    def shape(self):
        return self.h1.shape

    def float(self):
        """the fp32 value the planes represent (tests)"""
        return self.h1.float() + self.h2.float() / 2048.0


def tma_bn(n, rows):
    """column tile of the TMA-fed FP16-split kernel: the whole layer up to 256 columns; 128 when 256-column tiles would leave most SMs idle"""
    cap = int(_os.environ.get("EPNET_TMA_BN_MAX", "128"))  # 128: every tile double-buffers its accumulator (epilogue overlaps the MMAs)
    if n <= cap:
        return (n + 15) // 16 * 16
    bn = cap
    if cap == 256 and ((rows + 127) // 128) * ((n + 255) // 256) < 100 and n % 128 == 0:
        bn = 128
    return bn


class PackedLinear:
    """W (N,K) fp32 [+ bias (N)] -> the shared-memory image the MMA consumes.

    Per n-tile (BN rows) and k-block (32 columns) two planes, hi and lo, each BN x 128 bytes: K-major rows, grouped in
    8-row atoms of 1024 B, 16-byte chunk c of row r stored at chunk position c ^ (r % 8) (SWIZZLE_128B).
    hi = w with the 13 low mantissa bits cleared (exact TF32), lo = w - hi (exact in fp32)."""

    def __init__(self, weight, bias=None):
        w = weight.detach().float().contiguous()
        dev = w.device
        self.N, self.K = w.shape
        self.n_kblocks = (self.K + BK - 1) // BK
        self._w = w
        self._packs = {}
        # the FP16 split of the weights needs |w| < 65504 (h1 = fp16(w)); a layer outside that range stays on the TF32 split
        self.f16_ok = bool(w.numel() == 0 or (w.abs().max() < F16_LIMIT).item())
        self.BN = choose_bn(self.N)
        self.wpack = self._pack(self.BN)
        self.bias = None if bias is None else bias.detach().float().contiguous()

    def _pack(self, bn):
        if bn not in self._packs:
            w, dev = self._w, self._w.device
            n_tiles = (self.N + bn - 1) // bn
            npad, kpad = n_tiles * bn, self.n_kblocks * BK
            wp = torch.zeros(npad, kpad, device=dev)
            wp[:self.N, :self.K] = w
            hi = (wp.view(torch.int32) & -8192).view(torch.float32)  # 0xffffe000
            lo = wp - hi
            planes = torch.stack([hi, lo])                                    # (2, npad, kpad)
            planes = planes.view(2, n_tiles, bn, self.n_kblocks, 8, 4)          # k = kb*32 + c*4 + e
            planes = planes.permute(1, 3, 0, 2, 4, 5).contiguous()              # (tile, kb, plane, row, c, e)
            r = torch.arange(bn, device=dev) % 8
            c = torch.arange(8, device=dev)
            src_chunk = (c[None, :] ^ r[:, None])                               # stored position p holds chunk p ^ (r%8)
            idx = src_chunk[None, None, None, :, :, None].expand(n_tiles, self.n_kblocks, 2, bn, 8, 4)
            self._packs[bn] = torch.gather(planes, 4, idx).contiguous()
        return self._packs[bn]

    def _pack16(self, bn):
        """FP16 planes: h1 = fp16(w), h2 = fp16((w - h1) * 2^11); k-blocks of 64 halfs = one 128-byte swizzle row"""
        key = ("f16", bn)
        if key not in self._packs:
            w, dev = self._w, self._w.device
            n_tiles = (self.N + bn - 1) // bn
            nkb = (self.K + BK16 - 1) // BK16
            wp = torch.zeros(n_tiles * bn, nkb * BK16, device=dev)
            wp[:self.N, :self.K] = w
            h1 = wp.half()
            h2 = ((wp - h1.float()) * 2048.0).half()
            planes = torch.stack([h1, h2]).view(2, n_tiles, bn, nkb, 8, 8)       # k = kb*64 + c*8 + e
            planes = planes.permute(1, 3, 0, 2, 4, 5).contiguous()               # (tile, kb, plane, row, c, e)
            r = torch.arange(bn, device=dev) % 8
            c = torch.arange(8, device=dev)
            idx = (c[None, :] ^ r[:, None])[None, None, None, :, :, None].expand(n_tiles, nkb, 2, bn, 8, 8)
            self._packs[key] = torch.gather(planes, 4, idx).contiguous()
        return self._packs[key]

    def wide_f16(self, bn):
        """True when a launch with this column tile should take the FP16-split kernel"""
        return F16_WIDE and self.f16_ok and bn > 64

    def for_rows(self, rows):
        """(BN, packed weights) for a launch over `rows` rows: packed on first use per tile width, then cached (warm the shapes
        before capturing a CUDA graph)."""
        bn = choose_bn(self.N, rows)
        if self.K <= 128 and bn > 64:
            # at most 4 k-blocks per tile: the per-tile cost is set-up and epilogue, which the persistent narrow-tile kernel
            # amortises (measured: transposed conv 128 -> 16 x 4 x 4, 83 -> 58 us), even though X is then read once per 64 columns
            bn = choose_bn(min(self.N, 128))
        elif F16_WIDE and self.f16_ok and 64 < self.N <= 128 and self.K >= 512:
            # long-K layers with 65..128 columns (the 64->128 and 128->128 convolutions): one FP16-split tile instead of two narrow
            # TF32 tiles (235 -> 218 us, 133 -> 110 us); short-K layers of that width stay on the persistent narrow-tile kernel
            bn = (self.N + 15) // 16 * 16
        return bn, self._pack(bn)

    def from_planes(self, x, relu=True, pool=1, out=None, planes_out=False, f32_out=True):
        """x: Planes (L, >=K) -> (L / pool, N) fp32 and/or Planes: the TMA-fed FP16-split kernel (weights must be inside fp16's range)"""
        assert isinstance(x, Planes) and self.f16_ok and x.h1.dim() == 2 and x.h1.stride(-1) == 1
        L, ldx = x.shape[0], x.h1.stride(0)
        dev = x.h1.device
        if out is None and f32_out:
            out = torch.empty((L // pool, self.N), dtype=torch.float32, device=dev)
        planes = Planes.empty((L, self.N), dev) if planes_out else None
        bn = tma_bn(self.N, L)
        pc._call("gemm_planes_tma", LIB.epnet_gemm_planes_tma, x.h1, L, self.K, self.N, x.h1.data_ptr(), x.h2.data_ptr(), ldx,
                 self._pack16(bn).data_ptr(), bn, None if self.bias is None else self.bias.data_ptr(), int(bool(relu)), pool,
                 None if out is None else out.data_ptr(), 0 if out is None else out.stride(0),
                 None if planes is None else planes.h1.data_ptr(), None if planes is None else planes.h2.data_ptr(), self.N)
        return (out, planes) if planes_out else out

    def __call__(self, x, relu=True, pool=1, out=None, out_cm=None):
        """x (..., K) point-major rows (last dim contiguous) -> (rows / pool, N); with out_cm (B, N, pts) contiguous the result
        is written channel-major instead (rows = B * pts), straight from the epilogue."""
        assert x.is_cuda and x.dtype == torch.float32 and x.stride(-1) == 1
        x2 = x.reshape(-1, x.shape[-1]) if x.is_contiguous() else x
        assert x2.dim() == 2
        L, ldx = x2.shape[0], x2.stride(0)
        assert x2.shape[1] >= self.K
        if out_cm is not None:
            assert pool == 1 and out_cm.is_contiguous() and out_cm.shape[1] == self.N and out_cm.shape[0] * out_cm.shape[2] == L
            bn, wpack = self.for_rows(L)
            pc._call("gemm_tf32x3_cm", LIB.epnet_gemm_tf32x3_cm, x2, L, self.K, self.N, out_cm.shape[2], x2.data_ptr(), ldx, wpack.data_ptr(),
                     bn, None if self.bias is None else self.bias.data_ptr(), int(bool(relu)), out_cm.data_ptr())
            return out_cm
        if out is None:
            out = torch.empty((L // pool, self.N), dtype=torch.float32, device=x.device)
        assert out.stride(-1) == 1
        ldy = out.stride(0)
        bn, wpack = self.for_rows(L)
        fn = LIB.epnet_gemm_tf32x3
        if self.wide_f16(bn):
            fn, wpack = LIB.epnet_gemm_f16x3, self._pack16(bn)
        pc._call("gemm_tf32x3", fn, x2, L, self.K, self.N, x2.data_ptr(), ldx, wpack.data_ptr(), bn,
                 None if self.bias is None else self.bias.data_ptr(), int(bool(relu)), pool, out.data_ptr(), ldy)
        return out


def grouped_first_layer(lin, xyz, new_xyz, feats_pm, idx, relu=True, pool=1, out=None):
    """lin(group(xyz, new_xyz, feats, idx)) without the grouped tensor: lin = PackedLinear over [features (C) | offsets (3)];
    xyz (B,n,3), new_xyz (B,m,3), feats_pm (B,n,C) point-major or None, idx (B,m,ns) -> (B*m*ns / pool, N).
    Every width takes the fused path (tiles wider than 64 columns are cut into 64-column tiles of the narrow kernel)."""
    B, n = xyz.shape[0], xyz.shape[1]
    m, ns = idx.shape[1], idx.shape[2]
    c = 0 if feats_pm is None else feats_pm.shape[-1]
    assert lin.K == c + 3
    rows = B * m * ns
    bn, wpack = lin.for_rows(rows)
    if bn > 64:
        # the grouped operand exists on the narrow-tile kernel only (A through TMEM): wider layers (the 515 -> 256 first layers of the
        # last set-abstraction level) run as 64-column tiles of it, each gathering its rows again -- a few thousand rows, cheaper than
        # materialising QueryAndGroup's tensor and reading it back
        bn = 64
        wpack = lin._pack(bn)
    if out is None:
        out = torch.empty((rows // pool, lin.N), dtype=torch.float32, device=xyz.device)
    pc._call("gemm_tf32x3_grouped", LIB.epnet_gemm_tf32x3_grouped, xyz, B, n, m, ns, c, None if feats_pm is None else feats_pm.data_ptr(),
             0 if feats_pm is None else feats_pm.stride(-2), xyz.data_ptr(), new_xyz.data_ptr(), idx.data_ptr(), wpack.data_ptr(), bn, lin.N,
             None if lin.bias is None else lin.bias.data_ptr(), int(bool(relu)), pool, out.data_ptr(), out.stride(0))
    return out


class FusedFirstLevel:
    """One scale of a set-abstraction level WITHOUT input features in one kernel (csrc/sa_first_level.cu): group, re-centre, the
    three shared-MLP layers (BatchNorm folded) and the max over the ball, activations in registers, plain fp32 fma.  Built from the
    level's three PackedLinear layers; only the widths of the published configuration are instantiated."""
    SUPPORTED = {(16, 16, 32, 16), (32, 32, 64, 32)}  # (N1, N2, N3, nsample)

    @staticmethod
    def supports(lins, nsample):
        return (len(lins) == 3 and lins[0].K == 3 and lins[1].K == lins[0].N and lins[2].K == lins[1].N and
                (lins[0].N, lins[1].N, lins[2].N, nsample) in FusedFirstLevel.SUPPORTED)

    def __init__(self, lins, nsample):
        assert FusedFirstLevel.supports(lins, nsample)
        self.ns = nsample
        self.widths = (lins[0].N, lins[1].N, lins[2].N)
        dev = lins[0]._w.device
        bias = [lin.bias if lin.bias is not None else torch.zeros(lin.N, device=dev) for lin in lins]
        # layout of csrc/sa_first_level.cu SaPack: W1 rows (wx, wy, wz, bias) | W2 transposed (k-major) | b2 | W3 | b3
        parts = [torch.cat([lins[0]._w, bias[0][:, None]], dim=1).reshape(-1), lins[1]._w.t().contiguous().reshape(-1), bias[1],
                 lins[2]._w.reshape(-1), bias[2]]
        self.pack = torch.cat(parts).contiguous()

    def __call__(self, xyz, new_xyz, idx, out):
        """xyz (B,n,3), new_xyz (B,m,3), idx (B,m,ns) int32 -> out (B*m, N3) (row stride = out.stride(0), column stride 1)"""
        B, n, m = xyz.shape[0], xyz.shape[1], new_xyz.shape[1]
        assert idx.shape == (B, m, self.ns) and out.shape == (B * m, self.widths[2]) and out.stride(1) == 1
        pc._call("sa_first_level", LIB.epnet_sa_first_level, xyz, B, n, m, self.ns, *self.widths, xyz.data_ptr(), new_xyz.data_ptr(),
                 idx.data_ptr(), self.pack.data_ptr(), out.data_ptr(), out.stride(0))
        return out


class PackedDeconv:
    """ConvTranspose2d with kernel == stride on NHWC activations: one GEMM per map whose epilogue writes each input pixel's
    k x k x Cout patch into the output image (csrc/gemm_tf32x3.cu, dk mode).  weight (Cin, Cout, k, k) [+ bias (Cout)]."""

    def __init__(self, weight, bias=None):
        w = weight.detach().float()
        self.cin, self.cout, self.k = w.shape[0], w.shape[1], w.shape[2]
        assert w.shape[3] == self.k and self.cout % 4 == 0
        b = None if bias is None else bias.detach().float().repeat(self.k * self.k)  # column (ky, kx, o) gets bias[o]
        self.lin = PackedLinear(w.permute(2, 3, 1, 0).reshape(self.k * self.k * self.cout, self.cin), b)

    def __call__(self, x, out, relu=False):
        """x (B, h, w, >=Cin) NHWC (channel stride 1) -> out (B, h*k, w*k, Cout), a channel slice of an NHWC buffer"""
        assert x.is_cuda and x.dtype == torch.float32 and x.stride(-1) == 1 and x.shape[-1] >= self.cin
        B, h, w = x.shape[:3]
        ldx = x.stride(-2)
        assert x.stride(-3) == w * ldx and x.stride(0) == h * w * ldx
        ldo = out.stride(-2)
        assert out.shape == (B, h * self.k, w * self.k, self.cout) and out.stride(-1) == 1
        assert out.stride(-3) == w * self.k * ldo and out.stride(0) == h * self.k * w * self.k * ldo
        lin = self.lin
        bn, wpack = lin.for_rows(B * h * w)
        fn = LIB.epnet_deconv_nhwc_tf32x3
        if lin.wide_f16(bn):
            fn, wpack = LIB.epnet_deconv_nhwc_f16x3, lin._pack16(bn)
        pc._call("deconv_nhwc_tf32x3", fn, x, B, h, w, self.cin, self.k, self.cout, x.data_ptr(), ldx,
                 wpack.data_ptr(), bn, None if lin.bias is None else lin.bias.data_ptr(), int(bool(relu)), out.data_ptr(), ldo)
        return out


class PackedConv3x3:
    """3x3 / pad 1 convolution on NHWC activations as an implicit GEMM (csrc/gemm_tf32x3.cu, conv mode).
    weight (Cout, Cin, 3, 3) [+ bias]; Cin is zero-padded to a power of two >= 4 (the 3-channel input image becomes 4)."""

    def __init__(self, weight, bias=None, stride=1):
        w = weight.detach().float()
        cout, cin = w.shape[:2]
        cin_p = 4
        while cin_p < cin:
            cin_p *= 2
        wp = torch.zeros(cout, 3, 3, cin_p, device=w.device)
        wp[..., :cin] = w.permute(0, 2, 3, 1)
        self.lin = PackedLinear(wp.reshape(cout, 9 * cin_p), bias)
        self.cin, self.cin_p, self.cout, self.stride = cin, cin_p, cout, stride
        # 3 -> 64 channels, stride 1 (the first convolution of the image stream): K = 27 is too short for a tensor-core tile; a dedicated
        # fp32 FFMA kernel (csrc/first_conv.cu) takes the (o, ky, kx, c) weights as they are
        self.w_c3 = wp.contiguous() if (cin == 3 and cout == 64 and stride == 1) else None

    def planes_capable(self):
        """can this layer read its input as FP16 planes through TMA (k-blocks of one tap x 64 channels; weights inside fp16's range)?"""
        return F16_WIDE and self.lin.f16_ok and self.cin_p % 64 == 0

    def __call__(self, x, relu=False, out=None, planes_out=False, f32_out=True):
        """x (B, H, W, cin_p) NHWC contiguous fp32, or Planes of that shape -> (B, Ho, Wo, Cout) fp32.
        planes_out=True: also return the result as Planes for a following planes-fed layer -> (out, planes); with f32_out=False
        only the planes are written -> (None, planes)."""
        lin = self.lin
        if isinstance(x, Planes):
            assert self.planes_capable() and x.h1.is_cuda and x.shape[-1] >= self.cin_p and x.h1.stride(-1) == 1
            B, H, W, _ = x.shape
            ldx = x.h1.stride(-2)
            assert x.h1.stride(-3) == W * ldx and x.h1.stride(0) == H * W * ldx
        else:
            assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and x.shape[-1] == self.cin_p
            B, H, W, _ = x.shape
        Ho, Wo = (H - 1) // self.stride + 1, (W - 1) // self.stride + 1
        dev = x.h1.device if isinstance(x, Planes) else x.device
        if out is None and f32_out:
            out = torch.empty((B, Ho, Wo, self.cout), dtype=torch.float32, device=dev)
        planes = Planes.empty((B, Ho, Wo, self.cout), dev) if planes_out else None
        if out is not None:
            assert out.stride(-1) == 1 and out.stride(-3) == Wo * out.stride(-2) and out.stride(0) == Ho * out.stride(-3)
        ph1 = None if planes is None else planes.h1.data_ptr()
        ph2 = None if planes is None else planes.h2.data_ptr()
        if self.w_c3 is not None and not isinstance(x, Planes) and W % 4 == 0 and (out is None or (out.stride(-2) % 4 == 0 and out.data_ptr() % 16 == 0)):
            pc._call("conv3x3_c3_planes", LIB.epnet_conv3x3_c3_planes, x, B, H, W, self.cout, x.data_ptr(), self.w_c3.data_ptr(),
                     None if lin.bias is None else lin.bias.data_ptr(), int(bool(relu)), None if out is None else out.data_ptr(),
                     0 if out is None else out.stride(-2), ph1, ph2, self.cout)
            return (out, planes) if planes_out else out
        if isinstance(x, Planes):
            bn = tma_bn(self.cout, B * Ho * Wo)
            pc._call("conv3x3_planes_tma", LIB.epnet_conv3x3_planes_tma, x.h1, B, H, W, self.cin_p, self.cout, self.stride, x.h1.data_ptr(),
                     x.h2.data_ptr(), ldx, lin._pack16(bn).data_ptr(), bn, None if lin.bias is None else lin.bias.data_ptr(), int(bool(relu)),
                     None if out is None else out.data_ptr(), 0 if out is None else out.stride(-2), ph1, ph2, self.cout)
            return (out, planes) if planes_out else out
        if planes_out:  # fp32 input, planes (and optionally fp32) out: the TF32-split kernels
            bn = choose_bn(self.cout, B * Ho * Wo)
            pc._call("conv3x3_nhwc_tf32x3", LIB.epnet_conv3x3_nhwc_tf32x3_planes, x, B, H, W, self.cin_p, self.cout, self.stride, x.data_ptr(),
                     lin._pack(bn).data_ptr(), bn, None if lin.bias is None else lin.bias.data_ptr(), int(bool(relu)),
                     None if out is None else out.data_ptr(), 0 if out is None else out.stride(-2), ph1, ph2, self.cout)
            return out, planes
        bn, wpack = lin.for_rows(B * Ho * Wo)
        fn = LIB.epnet_conv3x3_nhwc_tf32x3
        if lin.wide_f16(bn):
            fn, wpack = LIB.epnet_conv3x3_nhwc_f16x3, lin._pack16(bn)
        pc._call("conv3x3_nhwc_tf32x3", fn, x, B, H, W, self.cin_p, self.cout, self.stride, x.data_ptr(),
                 wpack.data_ptr(), bn, None if lin.bias is None else lin.bias.data_ptr(), int(bool(relu)), out.data_ptr(),
                 out.stride(-2))
        return out
