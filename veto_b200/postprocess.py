"""Drop-in for ``PostProcessor.forward``, vanilla use_gt_box branch
(pysgg/modeling/roi_heads/relation_head/inference.py:398-453): softmax, max over predicate classes 1..,
triple score, per-image descending sort — one launch for the whole batch (``veto_postprocess``).

The reference's torch.sort is unstable; ties are broken by the original row (ascending).  The sgdet branch
(late NMS over boxes_per_cls, :414-432) and the MEET ensemble branches (:93-397) are not built yet.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops


class PostProcessor(nn.Module):
    def __init__(self, attribute_on=False, use_gt_box=True, later_nms_pred_thres=0.3, cfg=None):
        super().__init__()
        if attribute_on:
            raise NotImplementedError("attribute head is outside the VETO path")
        self.use_gt_box = use_gt_box
        self.later_nms_pred_thres = later_nms_pred_thres

    def forward(self, x, rel_pair_idxs, boxes):
        relation_logits, refine_logits = x
        if not self.use_gt_box:
            raise NotImplementedError("sgdet post-processing (obj_prediction_nms) is not built yet")
        if isinstance(relation_logits, dict):
            raise NotImplementedError("MEET ensemble post-processing is not built yet")
        n_boxes = [len(b) for b in boxes]
        rel_counts = [int(p.shape[0]) for p in rel_pair_idxs]
        obj_logit = torch.cat(list(refine_logits), 0)
        obj_prob = torch.softmax(obj_logit.float(), -1)
        obj_prob[:, 0] = 0  # :406
        obj_scores, obj_pred = obj_prob[:, 1:].max(dim=1)
        obj_pred = obj_pred + 1
        pairs_o, probs_o, labels_o, triple_o = ops.postprocess(torch.cat(list(relation_logits), 0),
                                                               torch.cat(list(rel_pair_idxs), 0), obj_scores,
                                                               rel_counts, n_boxes)
        results, ro, bo = [], 0, 0
        for box, nb, nr in zip(boxes, n_boxes, rel_counts):
            bl = box  # the reference adds the result fields to the input BoxList too (:431-452)
            bl.add_field("pred_labels", obj_pred[bo:bo + nb])
            bl.add_field("pred_scores", obj_scores[bo:bo + nb])
            bl.add_field("rel_pair_idxs", pairs_o[ro:ro + nr])
            bl.add_field("pred_rel_scores", probs_o[ro:ro + nr])
            bl.add_field("pred_rel_labels", labels_o[ro:ro + nr])
            bl.add_field("triple_scores", triple_o[ro:ro + nr])
            results.append(bl)
            ro += nr
            bo += nb
        return results


def make_roi_relation_post_processor(cfg):
    """inference.py:456-468."""
    return PostProcessor(False, cfg.MODEL.ROI_RELATION_HEAD.USE_GT_BOX, cfg.TEST.RELATION.LATER_NMS_PREDICTION_THRES, cfg)
