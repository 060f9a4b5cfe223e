"""Drop-in for ``PostProcessor.forward``, vanilla use_gt_box branch
(pysgg/modeling/roi_heads/relation_head/inference.py:398-453): softmax, max over predicate classes 1..,
triple score, per-image descending sort — one launch for the whole batch (``veto_postprocess``).

The reference's torch.sort is unstable; ties are broken by the original row (ascending).  In SGDet mode
(use_gt_box False) the object labels come from the late per-class NMS ``obj_prediction_nms`` (:414-417,
``veto_obj_nms_per_cls`` with late_nms) and the boxes are re-regressed to the chosen class (:425-431).
The MEET 'ensemble' branch (ENSEMBLE_LEARNING.ENABLED with EXPERT_GROUP False, :284-397) merges the group heads'
candidates of an image into one ranked list (``veto_postprocess_meet``); the reference only ever processes image 0
there (TEST.IMS_PER_BATCH 1), the drop-in handles the whole batch and is identical for a batch of one.  With
ENSEMBLE_LEARNING.EXPERT_GROUP the three experts of a group vote on every candidate (VOTING 'C' consensus / 'U'
unanimous, :93-283, ``veto_postprocess_meet_vote``).
"""
from __future__ import annotations

import itertools

import torch
import torch.nn as nn

from . import config as C
from . import ops
from .structures import BoxList


class PostProcessor(nn.Module):
    def __init__(self, attribute_on=False, use_gt_box=True, later_nms_pred_thres=0.3, cfg=None):
        super().__init__()
        if attribute_on:
            raise NotImplementedError("attribute head is outside the VETO path")
        self.use_gt_box = use_gt_box
        self.later_nms_pred_thres = later_nms_pred_thres
        self.cfg = cfg

    def forward(self, x, rel_pair_idxs, boxes, custom_rel_labels=None, cur_chosen_matrix=None, incre_idx_list=None,
                ensemble=False):
        relation_logits, refine_logits = x
        meet = isinstance(relation_logits, dict)
        vote = None
        if meet:
            if incre_idx_list is None:
                raise RuntimeError("MEET post-processing needs incre_idx_list (the predictor's 4th return value)")
            expert_group = self.cfg is not None and bool(C.get(self.cfg, "ENSEMBLE_LEARNING.EXPERT_GROUP", False))
            if expert_group:                                                        # :93-113: three experts per group
                voting = str(C.get(self.cfg, "ENSEMBLE_LEARNING.VOTING", "C"))
                if voting not in ("C", "U"):
                    raise RuntimeError("ENSEMBLE_LEARNING.VOTING must be 'C' (consensus) or 'U' (unanimous)")
                vote = voting == "C"
                n_groups = len(relation_logits) // 3
                names = ["group_%d%d" % (k, e) for e in (1, 2, 3) for k in range(n_groups)]
            else:
                n_groups = len(relation_logits)
                names = ["group_%d" % k for k in range(n_groups)]                  # :294-299
            head_sizes = [int(relation_logits[n].shape[1]) for n in names]
            col_map = []
            for i, n in enumerate(head_sizes):                                      # chosen_labels_incr (:351-353)
                k = i % n_groups
                members = [c for c, g in enumerate(incre_idx_list) if g == k + 1]
                if len(members) + 2 != n:
                    raise RuntimeError(f"head {names[i]} has {n} outputs but group {k} has {len(members)} member predicates")
                col_map += [0] + members + [0]                                      # last column (out of group) is dropped
            group_logits = torch.cat([relation_logits[n] for n in names], 1)
        n_boxes = [len(b) for b in boxes]
        rel_counts = [int(p.shape[0]) for p in rel_pair_idxs]
        obj_logit = torch.cat(list(refine_logits), 0)
        obj_prob = torch.softmax(obj_logit.float(), -1)
        obj_prob[:, 0] = 0  # :406
        boxes_per_cls = None
        if self.use_gt_box:
            obj_scores, obj_pred = obj_prob[:, 1:].max(dim=1)
            obj_pred = obj_pred + 1
        else:  # :414-417, late NMS for the object prediction
            boxes_per_cls = torch.cat([b.get_field("boxes_per_cls") for b in boxes], 0)
            obj_pred = ops.obj_nms_per_cls(obj_prob, boxes_per_cls, n_boxes, self.later_nms_pred_thres, late_nms=True)
            obj_scores = obj_prob.gather(1, obj_pred[:, None])[:, 0]
        row_starts = None
        if meet and vote is not None:
            pairs_o, probs_o, labels_o, triple_o, kept = ops.postprocess_meet_vote(
                group_logits, head_sizes, col_map, len(incre_idx_list), vote, torch.cat(list(rel_pair_idxs), 0), obj_scores,
                rel_counts, n_boxes)
            pairs_o = pairs_o.float()      # :272
            row_starts = [n_groups * o for o in itertools.accumulate([0] + rel_counts[:-1])]
            rel_counts = kept.tolist()     # survivors of the vote per image (one host sync)
        elif meet:
            G = len(head_sizes)
            pairs_o, probs_o, labels_o, triple_o = ops.postprocess_meet(
                group_logits, head_sizes, col_map, len(incre_idx_list), torch.cat(list(rel_pair_idxs), 0), obj_scores,
                rel_counts, n_boxes)
            pairs_o = pairs_o.float()      # the reference collects the merged pairs in a float32 tensor (:380)
            rel_counts = [G * r for r in rel_counts]
        else:
            pairs_o, probs_o, labels_o, triple_o = ops.postprocess(torch.cat(list(relation_logits), 0),
                                                                   torch.cat(list(rel_pair_idxs), 0), obj_scores,
                                                                   rel_counts, n_boxes)
        results, ro, bo = [], 0, 0
        for img, (box, nb, nr) in enumerate(zip(boxes, n_boxes, rel_counts)):
            if row_starts is not None:
                ro = row_starts[img]
            if self.use_gt_box:
                bl = box  # the reference adds the result fields to the input BoxList too (:431-452)
            else:         # sgdet: boxes regressed for the finetuned class (:425-431); a NEW BoxList without the input's fields
                cls = obj_pred[bo:bo + nb]
                # built with the INPUT's class: inside the reference that is pysgg's BoxList, whose resize() /
                # copy_with_fields() the reference evaluation calls on the result (vg_eval.py:55)
                bl = type(box)(boxes_per_cls[bo:bo + nb][torch.arange(nb, device=cls.device), cls], box.size, "xyxy")
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
