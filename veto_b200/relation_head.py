"""Drop-in for ``ROIRelationHead`` (pysgg/modeling/roi_heads/relation_head/relation_head.py:27-248), the caller of the
hot path, restricted to what the VETO predictors use: relation sampling (training) or candidate-pair enumeration
(test), the depth + RGB ROI gather, the predictor, the post-processor.  The relation-proposal network (``rel_pn``),
the attribute head, the union-feature extractor and BALANCED_NORM belong to other predictors and are refused.

    head = veto_b200.relation_head.build_roi_relation_head(cfg, in_channels)
    roi_features, result_or_proposals, losses = head(features, proposals, depth_features=depth, targets=targets)
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import config as C
from .postprocess import make_roi_relation_post_processor
from .registry import make_roi_box_feature_extractor, make_roi_relation_predictor
from .sampling import make_roi_relation_samp_processor


def to_onehot(labels: torch.Tensor, num_classes: int, fill: float = 1000.0) -> torch.Tensor:
    """model_kern.py:266-281: logits that softmax to a one-hot row (-fill everywhere, +fill at the label)."""
    out = torch.full((labels.shape[0], num_classes), -fill, dtype=torch.float32, device=labels.device)
    out[torch.arange(labels.shape[0], device=labels.device), labels.long()] = fill
    return out


class ROIRelationHead(nn.Module):
    def __init__(self, cfg, in_channels):
        super().__init__()
        self.cfg = cfg.clone() if hasattr(cfg, "clone") else cfg
        self.num_obj_cls, self.num_rel_cls = C.num_classes(cfg)                                  # :31-36
        rh = cfg.MODEL.ROI_RELATION_HEAD
        self.mode = ("predcls" if rh.USE_GT_OBJECT_LABEL else "sgcls") if rh.USE_GT_BOX else "sgdet"   # :39-45
        if rh.PREDICTOR not in ("VETOPredictor", "VETOPredictor_MEET"):
            raise NotImplementedError(f"veto_b200.ROIRelationHead serves the VETO predictors only, not {rh.PREDICTOR}")
        for key, what in (("MODEL.ROI_RELATION_HEAD.RELATION_PROPOSAL_MODEL.SET_ON", "the relation-proposal network"),
                          ("MODEL.ATTRIBUTE_ON", "the attribute head"), ("MODEL.BALANCED_NORM", "BALANCED_NORM")):
            if C.get(cfg, key, False):
                raise NotImplementedError(f"{what} is outside the VETO path")
        self.box_feature_extractor = make_roi_box_feature_extractor(cfg, in_channels, for_relation=True)    # :50-51
        self.predictor = make_roi_relation_predictor(cfg, 512)                                   # :52,64
        self.post_processor = make_roi_relation_post_processor(cfg)
        self.samp_processor = make_roi_relation_samp_processor(cfg)
        self.object_cls_refine = bool(C.get(cfg, "MODEL.ROI_RELATION_HEAD.OBJECT_CLASSIFICATION_REFINE", False))

    def forward(self, features, proposals, depth_features=None, targets=None, logger=None, x=None):
        """relation_head.py:90-248.  Training: (roi_features, proposals, losses); test: (roi_features, results, {})."""
        if self.mode == "predcls":                                                               # :104-111
            device = features[0].device
            for proposal in proposals:
                obj_labels = proposal.get_field("labels")
                proposal.add_field("predict_logits", to_onehot(obj_labels, self.num_obj_cls))
                proposal.add_field("pred_scores", torch.ones(len(obj_labels), device=device))
                proposal.add_field("pred_labels", obj_labels.to(device))
        if self.training:
            with torch.no_grad():                                                                # :112-132
                if self.cfg.MODEL.ROI_RELATION_HEAD.USE_GT_BOX:
                    proposals, rel_labels, rel_pair_idxs, _ = self.samp_processor.gtbox_relsample(proposals, targets)
                else:
                    proposals, rel_labels, _, rel_pair_idxs, _ = self.samp_processor.detect_relsample(proposals, targets)
        else:
            rel_labels = None
            rel_pair_idxs = self.samp_processor.prepare_test_pairs(features[0].device, proposals)  # :134-137
        roi_features, d_2d, _, _ = self.box_feature_extractor(features, proposals, depth_features=depth_features)  # :141
        obj_refine_logits, relation_logits, add_losses, incre_idx_list, cur_chosen_matrix, custom_rel_labels = self.predictor(
            proposals, rel_pair_idxs, rel_labels, logger, roi_features=roi_features, roi_depth_features=d_2d)   # :196-203
        if self.training:
            return roi_features, proposals, add_losses                                           # :247-248
        if not self.object_cls_refine:                                                           # :233-235
            obj_refine_logits = [p.get_field("predict_logits") for p in proposals]
        result = self.post_processor((relation_logits, obj_refine_logits), rel_pair_idxs, proposals,
                                     incre_idx_list=incre_idx_list, custom_rel_labels=custom_rel_labels,
                                     cur_chosen_matrix=cur_chosen_matrix,
                                     ensemble=bool(C.get(self.cfg, "ENSEMBLE_LEARNING.ENABLED", False)))   # :237-239
        return roi_features, result, {}


def build_roi_relation_head(cfg, in_channels):
    """relation_head.py:251-257."""
    return ROIRelationHead(cfg, in_channels)
