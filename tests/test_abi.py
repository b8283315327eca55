"""CPU-only checks of the drop-in boundary: libepnet_b200.so loads without a GPU and exports exactly the entry
points include/epnet_b200.h declares; argument validation works without launching anything."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "epnet_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(epnet_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from epnet_b200 import _lib
    names = _declared()
    assert len(names) >= 18
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), "missing export " + n
    assert set(_lib.SIGNATURES) | {"epnet_abi_version", "epnet_error_string"} == set(names)


def test_no_undeclared_exports():
    import subprocess
    from epnet_b200 import _lib
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted(l.split()[-1] for l in out.splitlines() if " T " in l and not l.split()[-1].startswith("_"))
    assert exported == _declared(), (exported, _declared())


def test_bad_arguments_are_reported_not_fatal():
    from epnet_b200 import _lib
    lib = _lib.LIB
    assert lib.epnet_abi_version() >= 1
    code = lib.epnet_furthest_point_sampling(1, 8, 2, None, None, None, None)
    assert code == -1
    assert b"bad argument" in lib.epnet_error_string(code)
    assert lib.epnet_error_string(0) == b"ok"
    # b == 0 is a no-op that must not touch the (null) stream or device
    assert lib.epnet_ball_query(0, 8, 2, 0.5, 4, 1, 1, 1, None) == 0
    assert lib.epnet_group_points(1, 0, 8, 2, 4, 1, 1, 1, None) == 0


def test_reference_function_names_present():
    # pointnet2_lib/pointnet2/src/pointnet2_api.cpp:11-23
    from epnet_b200 import pointnet2_cuda
    for n in ("ball_query_wrapper", "group_points_wrapper", "group_points_grad_wrapper", "gather_points_wrapper",
              "gather_points_grad_wrapper", "furthest_point_sampling_wrapper", "three_nn_wrapper",
              "three_interpolate_wrapper", "three_interpolate_grad_wrapper"):
        assert callable(getattr(pointnet2_cuda, n))


def test_install_registers_pointnet2_cuda():
    import sys
    import epnet_b200
    saved = sys.modules.get("pointnet2_cuda")
    try:
        epnet_b200.install()
        import pointnet2_cuda
        assert pointnet2_cuda is epnet_b200.pointnet2_cuda
    finally:
        if saved is None:
            sys.modules.pop("pointnet2_cuda", None)
        else:
            sys.modules["pointnet2_cuda"] = saved
