"""Drop-in for the registry-registered ``VETOFeatureExtractor``
(pysgg/modeling/roi_heads/box_head/roi_box_feature_extractors.py:75-141): ROIAlign 8x8 of the RGB FPN maps
(level per box) and of the depth map, returned as 2-D maps without FC layers.

The reference goes Pooler.forward -> LevelMapper + 4 x torch.nonzero + 5 ROIAlign launches
(pysgg/modeling/poolers.py:109-171); here the whole thing is one ``veto_roi_gather_forward`` launch.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import ops
from .registry import ROI_BOX_FEATURE_EXTRACTORS
from .structures import xyxy_boxes


class _RoiGather(torch.autograd.Function):
    """veto_roi_gather_forward with the backward the reference gets from _ROIAlign.backward
    (pysgg/layers/roi_align.py:26-44 -> _C.roi_align_backward): the gradient of the pooled depth features is
    scattered back into the depth map (the depth backbone is trainable, tools/relation_train_net.py:166-170).
    The RGB FPN maps come from the frozen backbone (relation_train_net.py:161-165) and get no gradient."""

    @staticmethod
    def forward(ctx, depth, boxes, n_boxes, scales, depth_scale, pool, sampling_ratio, k_min, k_max, *feats):
        x_2d, d_2d = ops.roi_gather(list(feats), depth, boxes, n_boxes, scales, depth_scale, pool=pool,
                                    sampling_ratio=sampling_ratio, k_min=k_min, k_max=k_max)
        ctx.save_for_backward(boxes)
        ctx.meta = (tuple(depth.shape), list(n_boxes), depth_scale, pool, sampling_ratio, len(feats))
        ctx.mark_non_differentiable(x_2d)
        return x_2d, d_2d

    @staticmethod
    def backward(ctx, g_x, g_d):
        (boxes,) = ctx.saved_tensors
        shape, n_boxes, depth_scale, pool, sampling_ratio, n_feats = ctx.meta
        g_depth = None
        if ctx.needs_input_grad[0] and g_d is not None:
            B, C, H, W = shape
            g_depth = ops.roi_align_backward(g_d.contiguous(), _rois(boxes, n_boxes), depth_scale, pool, pool, B, C, H, W,
                                             sampling_ratio)
        return (g_depth,) + (None,) * (8 + n_feats)


@ROI_BOX_FEATURE_EXTRACTORS.register("VETOFeatureExtractor")
class VETOFeatureExtractor(nn.Module):
    def __init__(self, cfg, in_channels, half_out=False, cat_all_levels=False, for_relation=False):
        super().__init__()
        if cat_all_levels:
            raise NotImplementedError("cat_all_levels is not used by the VETO path (relation_head.py:52-54)")
        self.resolution = cfg.MODEL.ROI_RELATION_HEAD.POOLER_RESOLUTION
        self.scales = tuple(cfg.MODEL.ROI_BOX_HEAD.POOLER_SCALES)
        self.sampling_ratio = cfg.MODEL.ROI_BOX_HEAD.POOLER_SAMPLING_RATIO
        # poolers.py:86-89: levels from the first / last scale
        self.k_min = int(round(-math.log2(self.scales[0])))
        self.k_max = int(round(-math.log2(self.scales[-1])))
        # the depth map is always pooled with the level-2 pooler when there are several (poolers.py:144-153)
        self.depth_scale = self.scales[2] if len(self.scales) > 1 else self.scales[0]
        self.out_channels = 256

    def forward(self, x, proposals, depth_features=None):
        if depth_features is None:
            raise NotImplementedError("the VETO path always passes depth_features (relation_head.py:141)")
        n_boxes = [len(p) for p in proposals]
        boxes = torch.cat([xyxy_boxes(p) for p in proposals], 0)
        feats = list(x)[:len(self.scales)]  # P6 is present in the reference's list but unused (poolers.py:157)
        if len(self.scales) == 1:
            x_2d = ops.roi_align_forward(feats[0], _rois(boxes, n_boxes), self.scales[0], self.resolution,
                                         self.resolution, self.sampling_ratio)
            d_2d = ops.roi_align_forward(depth_features, _rois(boxes, n_boxes), self.scales[0], self.resolution,
                                         self.resolution, self.sampling_ratio)
        else:
            if torch.is_grad_enabled() and depth_features.requires_grad:
                x_2d, d_2d = _RoiGather.apply(depth_features, boxes, n_boxes, self.scales, self.depth_scale,
                                              self.resolution, self.sampling_ratio, self.k_min, self.k_max, *feats)
            else:
                x_2d, d_2d = ops.roi_gather(feats, depth_features, boxes, n_boxes, self.scales, self.depth_scale,
                                            pool=self.resolution, sampling_ratio=self.sampling_ratio, k_min=self.k_min,
                                            k_max=self.k_max)
        return x_2d, d_2d, None, None


def _rois(boxes, n_boxes):
    """Pooler.convert_to_roi_format (poolers.py:96-107)."""
    ids = torch.repeat_interleave(torch.arange(len(n_boxes), device=boxes.device),
                                  torch.tensor(n_boxes, device=boxes.device), output_size=int(sum(n_boxes)))
    return torch.cat([ids[:, None].to(boxes.dtype), boxes], 1)
