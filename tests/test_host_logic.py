"""CPU-only checks of host-side decisions and of argument validation that happens before any CUDA call."""
import ctypes

import pytest


def test_choose_bn_properties():
    from epnet_b200.gemm import choose_bn, tile_policy
    for n in list(range(1, 300)) + [384, 512, 1000, 1024, 4096]:
        bn = choose_bn(n)
        tiles = (n + bn - 1) // bn
        assert bn % 16 == 0 and 16 <= bn <= 256
        assert tiles * bn >= n and (tiles - 1) * bn < n          # covers N with no empty tile
        if n <= 128:
            assert bn <= 64                                       # narrow tiles run on the A-from-TMEM kernel
    with tile_policy("latency"):
        few_rows = choose_bn(1024, rows=128)                      # one 128-row tile: cut the columns to fill SMs
        many_rows = choose_bn(1024, rows=1 << 20)
    with tile_policy("throughput"):
        wide = choose_bn(1024, rows=128)
    assert few_rows == 64 and many_rows == 256 and wide == 256


def test_tile_policy_restores_on_error():
    from epnet_b200 import gemm
    before = gemm._TILE_POLICY
    with pytest.raises(RuntimeError):
        with gemm.tile_policy("throughput"):
            assert gemm._TILE_POLICY == "throughput"
            raise RuntimeError("boom")
    assert gemm._TILE_POLICY == before
    with pytest.raises(AssertionError):
        with gemm.tile_policy("fastest"):
            pass


def test_new_entry_points_validate_before_launching():
    """every check below fails on arguments alone: no device is touched (this test runs on the CPU-only container)"""
    from epnet_b200 import _lib
    lib = _lib.LIB
    p = ctypes.c_void_p(4096)  # never dereferenced: validation comes first
    bad = _lib.EPNET_ERR_BAD_ARG if hasattr(_lib, "EPNET_ERR_BAD_ARG") else -1
    assert lib.epnet_bucket_cloud(1, 100, 100, p, p, p, None) == bad            # npad must be a power of two
    assert lib.epnet_bucket_cloud(1, 100, 32768, p, p, p, None) == bad          # more than one scene's shared memory
    assert lib.epnet_bucket_cloud(1, 200, 128, p, p, p, None) == bad            # npad < n
    assert lib.epnet_ball_query_sorted(1, 128, 8, 0.5, 65, p, p, p, p, None) == bad   # nsample > 64
    assert lib.epnet_ball_query_sorted(1, 100, 8, 0.5, 16, p, p, p, p, None) == bad   # npad not a power of two
    assert lib.epnet_gemm_tf32x3_grouped(1, 64, 8, 16, 4, p, 4, p, p, p, p, 128, 128, None, 1, 1, p, 128, None) == bad  # BN > 64
    assert lib.epnet_gemm_tf32x3_grouped(1, 64, 8, 16, 4, p, 2, p, p, p, p, 64, 64, None, 1, 1, p, 64, None) == bad     # ldf < c
    assert lib.epnet_gemm_tf32x3_cm(100, 8, 16, 30, p, 8, p, 16, None, 1, p, None) == bad                               # L % pts != 0
    assert lib.epnet_deconv_nhwc_tf32x3(1, 4, 4, 8, 2, 6, p, 8, p, 32, None, 0, p, 8, None) == bad                     # co % 4 != 0
    assert lib.epnet_attention_scale_pm(8, 6, 8, p, 8, p, 8, p, p, p, 8, p, 8, None) == bad                             # rc % 4 != 0


def _unswizzle(planes):
    """(tile, kb, plane, row, chunk position, e) -> logical chunk order: position p of row r holds chunk p ^ (r % 8)"""
    import torch
    t, kb, pl, rows, _, e = planes.shape
    r = torch.arange(rows) % 8
    p = torch.arange(8)
    idx = (p[None, :] ^ r[:, None])[None, None, None, :, :, None].expand(t, kb, pl, rows, 8, e)
    return torch.gather(planes, 4, idx)  # xor is its own inverse


def test_weight_packing_layouts_on_cpu():
    """The shared-memory images the GEMM kernels bulk-copy: TF32 hi/lo planes (k-blocks of 32 floats) and FP16 h1/h2 planes
    (k-blocks of 64 halfs), both K-major with the 128-byte swizzle.  Unswizzled, they must reproduce the weights: hi + lo exactly,
    h1 + 2^-11 h2 to 2^-21 relative; padding rows / columns are zero."""
    import torch
    from epnet_b200.gemm import PackedLinear
    g = torch.Generator().manual_seed(3)
    w = torch.randn(200, 99, generator=g)
    lin = PackedLinear(w, None)
    bn = lin.BN
    tiles = (200 + bn - 1) // bn
    planes = _unswizzle(lin._pack(bn))                                  # (tile, kb, 2, bn, 8, 4)
    full = planes.permute(2, 0, 3, 1, 4, 5).reshape(2, tiles * bn, -1)  # (plane, n, k)
    hi, lo = full[0], full[1]
    assert torch.equal(hi[:200, :99] + lo[:200, :99], w)
    assert torch.equal(hi.view(torch.int32) & 0x1FFF, torch.zeros_like(hi, dtype=torch.int32))  # hi is exactly TF32
    assert hi[200:].abs().max() == 0 and hi[:, 99:].abs().max() == 0 and lo[200:].abs().max() == 0
    p16 = _unswizzle(lin._pack16(bn))                                   # (tile, kb64, 2, bn, 8, 8) halfs
    full16 = p16.permute(2, 0, 3, 1, 4, 5).reshape(2, tiles * bn, -1).float()
    back = full16[0] + full16[1] / 2048.0
    assert (back[:200, :99] - w).abs().max() <= w.abs().max() * 2.0 ** -21
    assert back[200:].abs().max() == 0 and back[:, 99:].abs().max() == 0
    assert full16.shape[2] % 64 == 0 and full.shape[2] % 32 == 0


def test_round2_entry_points_validate_before_launching():
    """argument checks of the entry points added in round 2 (no device is touched)"""
    from epnet_b200 import _lib
    lib = _lib.LIB
    p = ctypes.c_void_p(4096)
    bad = _lib.EPNET_ERR_BAD_ARG if hasattr(_lib, "EPNET_ERR_BAD_ARG") else -1
    assert lib.epnet_sa_first_level(2, 16384, 4096, 32, 32, 32, 128, p, p, p, p, p, 128, None) == bad   # widths not instantiated
    assert lib.epnet_sa_first_level(2, 16384, 4096, 16, 32, 32, 64, p, p, p, p, p, 64, None) == bad     # nsample / width mismatch
    assert lib.epnet_sa_first_level(2, 16384, 4096, 32, 32, 32, 64, p, p, p, p, p, 32, None) == bad     # ldo < n3
    assert lib.epnet_sa_first_level(2, 16384, 4096, 32, 32, 32, 64, p, p, p, ctypes.c_void_p(4100), p, 64, None) == bad  # pack not 16-byte aligned
    assert lib.epnet_conv3x3_c3_planes(2, 384, 1282, 64, p, p, None, 1, p, 64, None, None, 0, None) == bad   # W % 4 != 0
    assert lib.epnet_conv3x3_c3_planes(2, 384, 1280, 32, p, p, None, 1, p, 64, None, None, 0, None) == bad   # cout != 64
    assert lib.epnet_conv3x3_c3_planes(2, 384, 1280, 64, p, p, None, 1, None, 0, None, None, 0, None) == bad  # no output at all
    assert lib.epnet_conv3x3_c3_planes(2, 384, 1280, 64, p, p, None, 1, None, 0, p, None, 64, None) == bad    # one plane without the other
    assert lib.epnet_fps_prefix_check(2, 1024, 2048, p, p, p, None) == bad                                    # m > n
    assert lib.epnet_fps_sample_guarded(2, 4096, 1024, p, p, p, p, None, None, 0, None, None) == bad          # no identity flags
    assert lib.epnet_fps_sample_guarded(2, 1024, 4096, p, p, p, p, None, None, 0, p, None) == bad             # m > n


def test_fused_first_level_pack_layout_on_cpu():
    """gemm.FusedFirstLevel packs the three folded layers for csrc/sa_first_level.cu (SaPack): W1 rows (wx, wy, wz, bias) | W2 transposed
    (k-major) | b2 | W3 | b3; supports() admits exactly the published first-level scales"""
    import torch
    from epnet_b200.gemm import FusedFirstLevel, PackedLinear
    g = torch.Generator().manual_seed(9)

    def layers(widths, first_k=3, bias=True):
        lins, k = [], first_k
        for n in widths:
            lins.append(PackedLinear(torch.randn(n, k, generator=g), torch.randn(n, generator=g) if bias else None))
            k = n
        return lins

    lins = layers((32, 32, 64))
    assert FusedFirstLevel.supports(lins, 32) and not FusedFirstLevel.supports(lins, 16)
    assert FusedFirstLevel.supports(layers((16, 16, 32)), 16)
    assert not FusedFirstLevel.supports(layers((32, 32, 64), first_k=4), 32)      # a level with an input feature (intensity)
    assert not FusedFirstLevel.supports(layers((64, 64, 128)), 32)                  # the second level's widths
    assert not FusedFirstLevel.supports(layers((32, 32)), 32)                       # two layers
    pack = FusedFirstLevel(lins, 32).pack
    n1, n2, n3 = 32, 32, 64
    assert pack.numel() == n1 * 4 + n1 * n2 + n2 + n3 * n2 + n3 and pack.data_ptr() % 16 == 0
    o = 0
    w1 = pack[o:o + n1 * 4].view(n1, 4); o += n1 * 4
    assert torch.equal(w1[:, :3], lins[0]._w) and torch.equal(w1[:, 3], lins[0].bias)
    w2t = pack[o:o + n1 * n2].view(n1, n2); o += n1 * n2
    assert torch.equal(w2t, lins[1]._w.t())
    assert torch.equal(pack[o:o + n2], lins[1].bias); o += n2
    assert torch.equal(pack[o:o + n3 * n2].view(n3, n2), lins[2]._w); o += n3 * n2
    assert torch.equal(pack[o:o + n3], lins[2].bias)
    nb = FusedFirstLevel(layers((16, 16, 32), bias=False), 16).pack  # layers without a bias: zeros in the bias slots
    assert nb[:64].view(16, 4)[:, 3].abs().max() == 0 and nb[64 + 256:64 + 256 + 16].abs().max() == 0
