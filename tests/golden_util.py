"""Loading of tests/golden/*.npz (made by tests/golden/make_golden.py from the reference's own Python layers)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name)))


def state_dict(g, prefix):
    return {k[len(prefix):]: torch.from_numpy(v) for k, v in g.items() if k.startswith(prefix)}


def small_backbone_config():
    from epnet_b200 import BackboneConfig
    return BackboneConfig(
        input_channels=0, npoints=[256, 64, 16, 4], radius=[[0.8, 2.0], [2.0, 4.0], [4.0, 8.0], [8.0, 16.0]],
        nsample=[[16, 32]] * 4,
        mlps=[[[8, 8, 16], [8, 8, 16]], [[16, 16, 32], [16, 24, 32]], [[32, 40, 48], [32, 40, 48]], [[48, 48, 64], [48, 56, 64]]],
        fp_mlps=[[32, 32], [48, 48], [64, 64], [64, 64]], li_fusion=True, image_attention=True, img_features_channel=32,
        img_channels=[3, 8, 16, 24, 32], point_channels=[32, 64, 96, 128], deconv_reduce=[4, 4, 4, 4])
