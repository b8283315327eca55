"""epnet_b200 -- B200 (sm_100a) implementation of the hot path of EPNet's two-stream RPN backbone:
the PointNet++ set-abstraction / feature-propagation ops and the LI-Fusion image-feature gather, behind
the reference's own `pointnet2_lib` op surface.

Importing the package loads libepnet_b200.so (building it with nvcc when missing); there is no CPU or
library fallback -- without the CUDA library the import fails.
"""
import sys

from . import _lib  # noqa: F401  (fails loudly when the CUDA library is unavailable)
from . import pointnet2_cuda, pointnet2_utils, pointnet2_modules, pytorch_utils, li_fusion  # noqa: F401
from .pointnet2_msg import BackboneConfig, Pointnet2MSG  # noqa: F401

__all__ = ["pointnet2_cuda", "pointnet2_utils", "pointnet2_modules", "pytorch_utils", "li_fusion", "BackboneConfig",
           "Pointnet2MSG", "install"]


def install(patch_grid_sample=True):
    """Make the UNCHANGED reference code run on these kernels.

    1. registers this package's `pointnet2_cuda` under that top-level name, which is what
       pointnet2_lib/pointnet2/pointnet2_utils.py:7 imports (likewise `iou3d_cuda` and `roipool3d_cuda`);
    2. rebinds the module-global `grid_sample` of the reference backbone (lib/net/pointnet2_msg.py:6, used
       at :118; lib/net/rpn.py:9 imports that file as top-level `pointnet2_msg`) if it is already
       imported, and returns a function that does so for modules imported later.
    """
    from . import iou3d_cuda, roipool3d_cuda
    sys.modules["pointnet2_cuda"] = pointnet2_cuda
    # the two next-row extensions: lib/utils/iou3d/iou3d_utils.py:2, lib/utils/roipool3d/roipool3d_utils.py:2
    sys.modules["iou3d_cuda"] = iou3d_cuda
    sys.modules["roipool3d_cuda"] = roipool3d_cuda

    def patch(module):
        module.grid_sample = li_fusion.grid_sample
        return module

    if patch_grid_sample:
        for name in ("pointnet2_msg", "lib.net.pointnet2_msg"):
            mod = sys.modules.get(name)
            if mod is not None and hasattr(mod, "Feature_Gather"):
                patch(mod)
    return patch
