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
