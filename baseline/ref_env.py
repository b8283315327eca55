"""Import the reference's own, UNMODIFIED Python (staged under baseline/_ref/ by baseline/stage_ref.py) in this
environment: torch 2.11, no `easydict`, and compiled extensions served by

    backend "reference": the reference's own CUDA kernels (oracle/_ref/libpointnet2_ref.so, built from the unmodified
                         sources by oracle/build_ref.sh), loaded WITHOUT importing the oracle package or epnet_b200 -- the
                         process of `bench.py --impl reference` maps libpointnet2_ref.so and nothing else of this repo;
    backend "product":   epnet_b200 (libepnet_b200.so) through epnet_b200.install().

Only the environment is adapted, never a reference file:
  * `easydict` is not installed -> a 20-line EasyDict (lib/config.py:1 needs attribute access and `type(v) is edict`);
  * lib/config.py:216 calls `yaml.load(f)` without a Loader, which PyYAML 6 refuses -> `load_yaml_cfg` reads the yaml itself
    and hands it to the reference's own `_merge_a_into_b` (lib/config.py:221-248);
  * lib/net/rpn.py:9 imports `pointnet2_msg` as a top-level module (tools/_init_path.py:5) -> baseline/_ref/lib/net is put on
    sys.path exactly like the reference's tools do.
Switching backend inside one process (tests compare both) rebinds the names the reference modules bound at import:
`pointnet2_utils.pointnet2`, `pointnet2_msg.grid_sample`, `iou3d_utils.iou3d_cuda`, `roipool3d_utils.roipool3d_cuda`.
"""
import importlib
import importlib.util
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(HERE, "_ref")
YAML = os.path.join(REF, "tools", "cfgs", "LI_Fusion_with_attention_use_ce_loss.yaml")


def staged():
    return os.path.exists(os.path.join(REF, "lib", "net", "pointnet2_msg.py"))


def _easydict_shim():
    if "easydict" in sys.modules:
        return

    class EasyDict(dict):
        def __init__(self, d=None, **kw):
            super().__init__()
            for k, v in dict(d or {}, **kw).items():
                setattr(self, k, v)

        def __setattr__(self, k, v):
            if isinstance(v, dict) and not isinstance(v, EasyDict):
                v = EasyDict(v)
            super().__setitem__(k, v)

        __setitem__ = __setattr__

        def __getattr__(self, k):
            try:
                return self[k]
            except KeyError:
                raise AttributeError(k)

    mod = types.ModuleType("easydict")
    mod.EasyDict = EasyDict
    sys.modules["easydict"] = mod


def load_by_path(name, path):
    """import a single file as module `name` without importing its package"""
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


_REF_KERNELS = None


def reference_kernels():
    """oracle/ref_cuda.py (ctypes doors onto oracle/_ref/libpointnet2_ref.so), loaded standalone."""
    global _REF_KERNELS
    if _REF_KERNELS is None:
        _REF_KERNELS = load_by_path("_epnet_ref_kernels", os.path.join(ROOT, "oracle", "ref_cuda.py"))
        if not _REF_KERNELS.available():
            raise RuntimeError("oracle/_ref/libpointnet2_ref.so is missing: run oracle/build_ref.sh where /root/reference exists")
    return _REF_KERNELS


def extension_modules(backend):
    """-> (pointnet2_cuda, iou3d_cuda, roipool3d_cuda, grid_sample) for a backend"""
    if backend == "reference":
        import torch.nn.functional as F
        k = reference_kernels()
        return k, k.iou3d_cuda, k.roipool3d_cuda, F.grid_sample
    if backend == "product":
        import epnet_b200
        from epnet_b200 import iou3d_cuda, li_fusion, pointnet2_cuda, roipool3d_cuda
        return pointnet2_cuda, iou3d_cuda, roipool3d_cuda, li_fusion.grid_sample
    raise ValueError(backend)


class Reference:
    """The imported reference modules plus `use(backend)`."""

    def __init__(self, mods):
        self.__dict__.update(mods)

    def use(self, backend):
        pn2, iou, roi, grid_sample = extension_modules(backend)
        self.pointnet2_utils.pointnet2 = pn2         # pointnet2_utils.py:7  `import pointnet2_cuda as pointnet2`
        self.pointnet2_msg.grid_sample = grid_sample  # lib/net/pointnet2_msg.py:6
        if "iou3d_utils" in self.__dict__:
            self.iou3d_utils.iou3d_cuda = iou          # lib/utils/iou3d/iou3d_utils.py:2
            self.roipool3d_utils.roipool3d_cuda = roi  # lib/utils/roipool3d/roipool3d_utils.py:2
        self.backend = backend
        return self


def load_yaml_cfg(cfg_module, path=YAML):
    import yaml
    from easydict import EasyDict
    with open(path) as f:
        cfg_module._merge_a_into_b(EasyDict(yaml.safe_load(f)), cfg_module.cfg)


def import_reference(backend="reference", with_rcnn=False, yaml_cfg=True):
    """-> Reference(pointnet2_utils, pointnet2_modules, pytorch_utils, pointnet2_msg, cfg[, point_rcnn, iou3d_utils,
    roipool3d_utils, proposal_layer]) bound to `backend`."""
    if not staged():
        raise RuntimeError("baseline/_ref is empty: run `python -m baseline.stage_ref` where /root/reference exists")
    _easydict_shim()
    pn2, iou, roi, _ = extension_modules(backend)
    sys.modules["pointnet2_cuda"] = pn2
    sys.modules["iou3d_cuda"] = iou
    sys.modules["roipool3d_cuda"] = roi
    for p in (os.path.join(REF, "lib", "net"), REF):  # tools/_init_path.py:3-5
        if p not in sys.path:
            sys.path.insert(0, p)
    import lib.config as config
    first = not getattr(config, "_epnet_yaml_loaded", False)
    if yaml_cfg and first:
        load_yaml_cfg(config)
        config._epnet_yaml_loaded = True
    from pointnet2_lib.pointnet2 import pointnet2_modules, pointnet2_utils, pytorch_utils
    pointnet2_msg = importlib.import_module("pointnet2_msg")  # the module object lib/net/rpn.py:9 uses
    mods = dict(pointnet2_utils=pointnet2_utils, pointnet2_modules=pointnet2_modules, pytorch_utils=pytorch_utils,
                pointnet2_msg=pointnet2_msg, cfg=config.cfg, config=config)
    if with_rcnn:
        import lib.utils.iou3d.iou3d_utils as iou3d_utils
        import lib.utils.roipool3d.roipool3d_utils as roipool3d_utils
        import lib.rpn.proposal_layer as proposal_layer
        point_rcnn = importlib.import_module("lib.net.point_rcnn")
        mods.update(iou3d_utils=iou3d_utils, roipool3d_utils=roipool3d_utils, proposal_layer=proposal_layer, point_rcnn=point_rcnn)
    return Reference(mods).use(backend)
