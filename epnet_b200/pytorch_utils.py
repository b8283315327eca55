"""Shared-MLP building blocks with the module tree of the reference's
pointnet2_lib/pointnet2/pytorch_utils.py, so published EPNet/PointRCNN checkpoints load key for key:
`<mlp>.layer{i}.conv.weight`, `<mlp>.layer{i}.bn.bn.{weight,bias,running_mean,running_var}`.

Only what the hot path instantiates is provided: SharedMLP (pytorch_utils.py:5-32) over 1x1 Conv2d
units (:163-199) and the Conv1d unit (:126-160) used by the RPN/RCNN heads.
"""
from typing import List

import torch.nn as nn


class _NormHolder(nn.Sequential):
    """BatchNorm wrapped in a Sequential under the child name 'bn' (pytorch_utils.py:104-111)."""

    def __init__(self, channels: int, norm_cls, name: str = ""):
        super().__init__()
        norm = norm_cls(channels)
        nn.init.constant_(norm.weight, 1.0)
        nn.init.constant_(norm.bias, 0)
        self.add_module(name + "bn", norm)


class BatchNorm1d(_NormHolder):
    def __init__(self, in_size: int, *, name: str = ""):
        super().__init__(in_size, nn.BatchNorm1d, name)


class BatchNorm2d(_NormHolder):
    def __init__(self, in_size: int, name: str = ""):
        super().__init__(in_size, nn.BatchNorm2d, name)


class _ConvUnit(nn.Sequential):
    """conv -> [bn] -> [activation]  (or the pre-activation order), cf. pytorch_utils.py:35-101."""

    def __init__(self, conv_cls, norm_holder_cls, inorm_cls, in_size, out_size, *, kernel_size, stride, padding,
                 activation, bn, init, bias, preact, name, instance_norm):
        super().__init__()
        use_bias = bias and not bn
        conv = conv_cls(in_size, out_size, kernel_size=kernel_size, stride=stride, padding=padding, bias=use_bias)
        init(conv.weight)
        if use_bias:
            nn.init.constant_(conv.bias, 0)
        norm_channels = in_size if preact else out_size

        def add_post_ops():
            if bn:
                self.add_module(name + "bn", norm_holder_cls(norm_channels))
            if activation is not None:
                self.add_module(name + "activation", activation)
            if not bn and instance_norm:
                self.add_module(name + "in", inorm_cls(norm_channels, affine=False, track_running_stats=False))

        if preact:
            add_post_ops()
        self.add_module(name + "conv", conv)
        if not preact:
            add_post_ops()


class Conv1d(_ConvUnit):
    def __init__(self, in_size: int, out_size: int, *, kernel_size: int = 1, stride: int = 1, padding: int = 0,
                 activation=nn.ReLU(inplace=True), bn: bool = False, init=nn.init.kaiming_normal_, bias: bool = True,
                 preact: bool = False, name: str = "", instance_norm=False):
        super().__init__(nn.Conv1d, BatchNorm1d, nn.InstanceNorm1d, in_size, out_size, kernel_size=kernel_size,
                         stride=stride, padding=padding, activation=activation, bn=bn, init=init, bias=bias,
                         preact=preact, name=name, instance_norm=instance_norm)


class Conv2d(_ConvUnit):
    def __init__(self, in_size: int, out_size: int, *, kernel_size=(1, 1), stride=(1, 1), padding=(0, 0),
                 activation=nn.ReLU(inplace=True), bn: bool = False, init=nn.init.kaiming_normal_, bias: bool = True,
                 preact: bool = False, name: str = "", instance_norm=False):
        super().__init__(nn.Conv2d, BatchNorm2d, nn.InstanceNorm2d, in_size, out_size, kernel_size=kernel_size,
                         stride=stride, padding=padding, activation=activation, bn=bn, init=init, bias=bias,
                         preact=preact, name=name, instance_norm=instance_norm)


class SharedMLP(nn.Sequential):
    """Stack of 1x1 Conv2d units named layer0, layer1, ... (pytorch_utils.py:5-32)."""

    def __init__(self, args: List[int], *, bn: bool = False, activation=nn.ReLU(inplace=True), preact: bool = False,
                 first: bool = False, name: str = "", instance_norm: bool = False):
        super().__init__()
        for i in range(len(args) - 1):
            plain = first and preact and i == 0  # the very first pre-activation layer has no bn/activation
            self.add_module(
                name + "layer{}".format(i),
                Conv2d(args[i], args[i + 1], bn=bn and not plain, activation=None if plain else activation,
                       preact=preact, instance_norm=instance_norm))
