"""Drop-in for the depth backbone: ``registry.BACKBONES["R-18-C4"]`` = ``build_resnet18_depth``
(pysgg/modeling/backbone/backbone.py:83-93), i.e. ``nn.Sequential(body=ResNetDepth)`` with ``out_channels = 256``
(pysgg/modeling/backbone/resnet_depth.py:11-47: torchvision's ResNet-18, one-channel conv1, truncated after layer3).

Same state-dict keys and order as the reference's module (``body.conv1.weight``, ``body.bn1.*``,
``body.layer2.0.downsample.1.running_var`` …), same initialisation, same train()/eval() BatchNorm behaviour.  The
whole forward is one C call (``veto_depth_backbone_forward``) and the whole backward another; there is no PyTorch
fallback.  ``detector/generalized_rcnn.py:53-54`` calls it as ``self.depth_backbone(depth_images.tensors)``.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch
import torch.nn as nn

from . import config as C
from . import ops
from .registry import BACKBONES


class _Conv(nn.Module):
    """nn.Conv2d(cin, cout, k, bias=False) as a parameter holder (torchvision/models/resnet.py conv3x3 / conv1x1)."""

    def __init__(self, cin, cout, k):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(cout, cin, k, k))
        nn.init.kaiming_normal_(self.weight, mode="fan_out", nonlinearity="relu")   # ResNet.__init__


class _BN(nn.Module):
    """nn.BatchNorm2d(c) as a parameter / buffer holder (momentum 0.1, eps 1e-5)."""

    def __init__(self, c):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(c))
        self.bias = nn.Parameter(torch.zeros(c))
        self.register_buffer("running_mean", torch.zeros(c))
        self.register_buffer("running_var", torch.ones(c))
        self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))


class _Block(nn.Module):
    """BasicBlock: conv1, bn1, conv2, bn2[, downsample = (conv1x1, bn)]."""

    def __init__(self, cin, cout, stride):
        super().__init__()
        self.conv1, self.bn1 = _Conv(cin, cout, 3), _BN(cout)
        self.conv2, self.bn2 = _Conv(cout, cout, 3), _BN(cout)
        if stride != 1 or cin != cout:
            self.downsample = nn.Sequential(_Conv(cin, cout, 1), _BN(cout))

    def pairs(self):
        out = [(self.conv1, self.bn1), (self.conv2, self.bn2)]
        if hasattr(self, "downsample"):
            out.append((self.downsample[0], self.downsample[1]))
        return out


class _DepthFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, depth, *params):
        convs, bn_w, bn_b, means, vars_ = module._groups()
        out, ws = ops.depth_backbone_forward(depth, convs, bn_w, bn_b, means, vars_, True, 0.1, module.precision)
        ctx.module, ctx.ws, ctx.shape = module, ws, tuple(depth.shape)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        m = ctx.module
        if ctx.ws is None:
            raise RuntimeError("the depth backbone's saved activations were already released (backward called twice)")
        convs, bn_w, bn_b, means, vars_ = m._groups()
        _, g_conv, g_w, g_b = ops.depth_backbone_backward(grad_out, ctx.shape, convs, bn_w, bn_b, means, vars_, ctx.ws,
                                                          m.precision)
        ctx.ws = None
        return (None, None) + tuple(g_conv) + tuple(g_w) + tuple(g_b)


class ResNetDepth(nn.Module):
    def __init__(self, precision="bf16x3"):
        super().__init__()
        self.precision = precision
        self.conv1 = _Conv(1, 64, 7)
        n = 7 * 7 * 64
        self.conv1.weight.data.normal_(0, math.sqrt(2.0 / n))                       # resnet_depth.py:27-28
        self.bn1 = _BN(64)
        self.layer1 = nn.Sequential(_Block(64, 64, 1), _Block(64, 64, 1))
        self.layer2 = nn.Sequential(_Block(64, 128, 2), _Block(128, 128, 1))
        self.layer3 = nn.Sequential(_Block(128, 256, 2), _Block(256, 256, 1))

    def _pairs(self):
        """(conv, bn) holders in the module order of include/veto_b200.h."""
        out = [(self.conv1, self.bn1)]
        for layer in (self.layer1, self.layer2, self.layer3):
            for block in layer:
                out += block.pairs()
        return out

    def _groups(self):
        p = self._pairs()
        return ([c.weight for c, _ in p], [b.weight for _, b in p], [b.bias for _, b in p],
                [b.running_mean for _, b in p], [b.running_var for _, b in p])

    def forward(self, x):
        if self.training:
            convs, bn_w, bn_b, _, _ = self._groups()
            out = _DepthFn.apply(self, x, *convs, *bn_w, *bn_b)
            for _, b in self._pairs():
                b.num_batches_tracked += 1
            return out
        # eval(): BatchNorm on the running statistics; inference only (the reference never back-propagates through an
        # eval-mode depth backbone: relation_train_net.py:166-170 keeps it in train_modules)
        convs, bn_w, bn_b, means, vars_ = self._groups()
        with torch.no_grad():
            out, _ = ops.depth_backbone_forward(x, convs, bn_w, bn_b, means, vars_, False, 0.1, self.precision)
        return out


@BACKBONES.register("R-18-C4")
def build_resnet18_depth(cfg, depth_backbone=False):
    """backbone.py:83-93."""
    body = ResNetDepth(C.get(cfg, "VETO_B200.PRECISION", "f16c8"))
    model = nn.Sequential(OrderedDict([("body", body)]))
    model.out_channels = 256
    return model


def build_backbone(cfg, depth_backbone=False):
    """backbone.py:95-106, the depth branch (the RGB detector backbone is outside this library)."""
    if not depth_backbone:
        raise NotImplementedError("only the depth backbone (MODEL.DEPTH_BACKBONE.CONV_BODY) is part of the VETO path")
    name = C.get(cfg, "MODEL.DEPTH_BACKBONE.CONV_BODY", "R-18-C4")
    assert name in BACKBONES, "cfg.MODEL.DEPTH_BACKBONE.CONV_BODY: {} are not registered in registry".format(name)
    return BACKBONES[name](cfg, depth_backbone)
