"""Drop-in for ``RelationSampling.prepare_test_pairs``
(pysgg/modeling/roi_heads/relation_head/sampling.py:31-52): candidate-pair enumeration on the device in one
launch for the whole batch, instead of a per-image loop of ones/eye/nonzero/sort.

Tie-breaking over the MAX_PROPOSAL_PAIR cap: the reference uses an unstable torch.sort on
pred_scores[i]*pred_scores[j]; every pair ties with its mirror.  Here the order is defined as
(product descending, row-major pair index ascending) — what the reference's CPU path produces.
"""
from __future__ import annotations

import torch

from . import ops
from .structures import xyxy_boxes


class RelationSampling:
    def __init__(self, fg_thres=0.5, require_overlap=False, num_sample_per_gt_rel=4, batch_size_per_image=1024,
                 positive_fraction=0.25, max_proposal_pairs=2048, use_gt_box=True, test_overlap=False):
        self.fg_thres = fg_thres
        self.require_overlap = require_overlap
        self.num_sample_per_gt_rel = num_sample_per_gt_rel
        self.batch_size_per_image = batch_size_per_image
        self.positive_fraction = positive_fraction
        self.use_gt_box = use_gt_box
        self.max_proposal_pairs = max_proposal_pairs
        self.test_overlap = test_overlap

    def prepare_test_pairs(self, device, proposals):
        n_boxes = [len(p) for p in proposals]
        overlap = (not self.use_gt_box) and self.test_overlap
        over_cap = any(n * (n - 1) > self.max_proposal_pairs for n in n_boxes)
        boxes = torch.cat([xyxy_boxes(p) for p in proposals], 0).to(device) if overlap else None
        scores = torch.cat([p.get_field("pred_scores") for p in proposals], 0).to(device) if over_cap else None
        return ops.enumerate_pairs(n_boxes, device, self.max_proposal_pairs, boxes=boxes, scores=scores,
                                   require_overlap=overlap)

    def gtbox_relsample(self, proposals, targets):
        raise NotImplementedError("training-time samplers are a 'next' row (SURVEY.md §8 f2)")

    def detect_relsample(self, proposals, targets):
        raise NotImplementedError("training-time samplers are a 'next' row (SURVEY.md §8 f2)")


def make_roi_relation_samp_processor(cfg):
    """sampling.py:312-324."""
    rh = cfg.MODEL.ROI_RELATION_HEAD
    return RelationSampling(
        max_proposal_pairs=rh.MAX_PROPOSAL_PAIR,
        use_gt_box=rh.USE_GT_BOX,
        test_overlap=cfg.TEST.RELATION.REQUIRE_OVERLAP,
    )
