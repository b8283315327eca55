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
