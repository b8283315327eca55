"""Stage the reference's own Python files into baseline/_ref/ (git-ignored; ships to the GPU box with the snapshot).

    python -m baseline.stage_ref            # needs /root/reference

Copies, byte for byte: pointnet2_lib/pointnet2/*.py, lib/**/*.py (network, config, proposal layers, utils) and
tools/cfgs/*.yaml.  No CUDA/C++ sources are copied (the kernels are compiled where they lie by oracle/build_ref.sh) and
nothing is modified: baseline/ref_env.py adapts the ENVIRONMENT (easydict shim, extension modules), never the files.
On the GPU box /root/reference does not exist and this is a no-op that keeps the staged copy."""
import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SRC = os.environ.get("EPNET_REFERENCE_ROOT", "/root/reference")

GROUPS = [("pointnet2_lib/pointnet2", (".py",), False), ("lib", (".py",), True), ("tools/cfgs", (".yaml",), False)]
SKIP = {"setup.py"}


def _files():
    for rel, exts, recursive in GROUPS:
        top = os.path.join(SRC, rel)
        for dirpath, dirnames, filenames in os.walk(top):
            dirnames[:] = [d for d in dirnames if d not in ("__pycache__", "src", "build")] if recursive else []
            for f in sorted(filenames):
                if f.endswith(exts) and f not in SKIP:
                    yield os.path.relpath(os.path.join(dirpath, f), SRC)


def stage(verbose=False):
    """-> number of files staged (0 when /root/reference is absent)."""
    if not os.path.isdir(SRC):
        if verbose:
            print("stage_ref: %s not present (GPU box?) -- keeping %s" % (SRC, DEST))
        return 0
    n = 0
    for rel in _files():
        dst = os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not (os.path.exists(dst) and filecmp.cmp(os.path.join(SRC, rel), dst, shallow=False)):
            shutil.copyfile(os.path.join(SRC, rel), dst)
        n += 1
    if verbose:
        print("stage_ref: %d reference files under %s" % (n, DEST))
    return n


def staged():
    return os.path.exists(os.path.join(DEST, "lib", "net", "pointnet2_msg.py"))


def verify():
    """every staged file equals its source (only meaningful where /root/reference exists) -> list of differing files"""
    return [rel for rel in _files() if not filecmp.cmp(os.path.join(SRC, rel), os.path.join(DEST, rel), shallow=False)]


if __name__ == "__main__":
    stage(verbose=True)
    sys.exit(0)
