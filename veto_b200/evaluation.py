"""Recall of the scene-graph evaluation on the device.

Mirrors ``SGRecall.calculate_recall`` (pysgg/data/datasets/evaluation/vg/sgg_eval.py:138-186) for a batch of images:
the reference moves every prediction to the host, builds (subject class, predicate, object class) triplets with numpy
and matches them image by image on rank 0; here the predictions stay where the post-processor left them and
``veto_sgg_match`` matches all images in one launch.  Only the final per-image numbers cross to the host.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import torch

from . import ops


def triplets(rel_pairs: torch.Tensor, rel_labels: torch.Tensor, classes: torch.Tensor, boxes: torch.Tensor):
    """_triplet (sgg_eval.py:44-74): (triplets int64 [R,3] = (subject class, predicate, object class), boxes [R,8])."""
    s, o = rel_pairs[:, 0].long(), rel_pairs[:, 1].long()
    return torch.stack((classes[s].long(), rel_labels.long(), classes[o].long()), 1), torch.cat((boxes[s], boxes[o]), 1)


def mean_recall(first_match: Sequence[torch.Tensor], gt_predicates: Sequence[torch.Tensor], num_rel: int,
                ks: Sequence[int] = (20, 50, 100)) -> Dict[str, object]:
    """SGMeanRecall.collect_mean_recall_items + calculate_mean_recall (sgg_eval.py:424-466) from the first-match ranks:
    per image and predicate class the fraction of its ground-truth triplets matched within the first k predictions;
    a class's recall = the mean over the images that contain it (0 if none does); mR@k = the mean over the num_rel - 1
    foreground classes.  Returns {'mean_recall': {k: float}, 'mean_recall_list': {k: [per class]}}."""
    collect = {k: [[] for _ in range(num_rel)] for k in ks}
    for fm, pr in zip(first_match, gt_predicates):
        pr = pr.long().cpu()
        count = torch.bincount(pr, minlength=num_rel)
        for k in ks:
            hit = torch.bincount(pr[fm.cpu() < k], minlength=num_rel)
            for n in torch.nonzero(count).flatten().tolist():
                collect[k][n].append(float(hit[n]) / float(count[n]))
    out = {"mean_recall": {}, "mean_recall_list": {}}
    for k in ks:
        per_class = [sum(v) / len(v) if v else 0.0 for v in collect[k][1:]]
        out["mean_recall_list"][k] = per_class
        out["mean_recall"][k] = sum(per_class) / float(num_rel - 1)
    return out


def recall_at_k(predictions, groundtruths, ks: Sequence[int] = (20, 50, 100), iou_thres: float = 0.5,
                predcls_like: bool = False) -> Dict[str, object]:
    """predictions: the post-processor's BoxLists (fields rel_pair_idxs, pred_rel_scores, pred_labels; ranked rows);
    groundtruths: BoxLists with fields labels and relation_tuple [G,3] = (subject idx, object idx, predicate)
    (vg_eval.py:469-495).  predcls_like: take boxes / classes from the ground truth (PredCls, vg_eval.py:512-515).
    Returns {'recall': {k: [per image]}, 'hits_per_rel': {k: {predicate: [hits, count]}}, 'first_match': [per image],
    'gt_predicates': [per image]} — feed the last two to mean_recall() for mR@k.
    Images without ground-truth relations are skipped like the reference does (vg_eval.py:473-474)."""
    gt_t, gt_b, pr_t, pr_b, gt_n, pr_n, gt_pred = [], [], [], [], [], [], []
    for pred, gt in zip(predictions, groundtruths):
        rel_tuple = gt.get_field("relation_tuple").long()
        if rel_tuple.shape[0] == 0:
            continue
        gt_cls, gt_box = gt.get_field("labels").long(), gt.convert("xyxy").bbox
        t, b = triplets(rel_tuple[:, :2], rel_tuple[:, 2], gt_cls, gt_box)
        gt_t.append(t)
        gt_b.append(b)
        gt_n.append(t.shape[0])
        gt_pred.append(rel_tuple[:, 2])
        scores = pred.get_field("pred_rel_scores")
        labels = 1 + scores[:, 1:].argmax(1)                                       # :152
        cls, box = (gt_cls, gt_box) if predcls_like else (pred.get_field("pred_labels").long(), pred.convert("xyxy").bbox)
        t, b = triplets(pred.get_field("rel_pair_idxs").long(), labels, cls, box)
        pr_t.append(t)
        pr_b.append(b)
        pr_n.append(t.shape[0])
    out = {"recall": {k: [] for k in ks}, "hits_per_rel": {k: {} for k in ks}, "first_match": [], "gt_predicates": []}
    if not gt_t:
        return out
    first, _ = ops.sgg_match(torch.cat(gt_t), torch.cat(gt_b), gt_n, torch.cat(pr_t), torch.cat(pr_b), pr_n, iou_thres)
    first_host = first.cpu()
    preds_host = torch.cat(gt_pred).cpu()
    off = 0
    for n in gt_n:
        fm, pr = first_host[off:off + n], preds_host[off:off + n]
        off += n
        out["first_match"].append(fm)
        out["gt_predicates"].append(pr)
        for k in ks:
            hit = fm < k
            out["recall"][k].append(float(hit.sum()) / float(n))                    # :158-160
            per = out["hits_per_rel"][k]
            for r, h in zip(pr.tolist(), hit.tolist()):                             # :161-168
                e = per.setdefault(r, [0, 0])
                e[0] += int(h)
                e[1] += 1
    return out


def recall_nogc_at_k(predictions, groundtruths, ks: Sequence[int] = (20, 50, 100), iou_thres: float = 0.5,
                     predcls_like: bool = False, top: int = 100) -> Dict[str, object]:
    """SGNoGraphConstraintRecall.calculate_recall (sgg_eval.py:213-252): every (pair, predicate) combination is a
    candidate, scored obj_s * obj_o * rel_scores[pair, predicate]; the `top` best per image (descending, ties by flat
    index — the reference's numpy argsort is not stable) are matched like the graph-constrained ones.
    Returns {'recall': {k: [per image]}, 'first_match': [per image]}."""
    gt_t, gt_b, pr_t, pr_b, gt_n, pr_n = [], [], [], [], [], []
    for pred, gt in zip(predictions, groundtruths):
        rel_tuple = gt.get_field("relation_tuple").long()
        if rel_tuple.shape[0] == 0:
            continue
        gt_cls, gt_box = gt.get_field("labels").long(), gt.convert("xyxy").bbox
        t, b = triplets(rel_tuple[:, :2], rel_tuple[:, 2], gt_cls, gt_box)
        gt_t.append(t)
        gt_b.append(b)
        gt_n.append(t.shape[0])
        scores = pred.get_field("pred_rel_scores")
        pairs = pred.get_field("rel_pair_idxs").long()
        if predcls_like:
            cls, box = gt_cls, gt_box
            obj_scores = torch.ones(len(gt_cls), dtype=scores.dtype, device=scores.device)     # vg_eval.py:515
        else:
            cls, box = pred.get_field("pred_labels").long(), pred.convert("xyxy").bbox
            obj_scores = pred.get_field("pred_scores")
        overall = (obj_scores[pairs[:, 0]] * obj_scores[pairs[:, 1]])[:, None] * scores[:, 1:]   # :221-222
        order = torch.sort(overall.reshape(-1), descending=True, stable=True)[1][:top]            # :223
        row, col = order // overall.shape[1], order % overall.shape[1]
        t, b = triplets(pairs[row], col + 1, cls, box)                                            # :224-231
        pr_t.append(t)
        pr_b.append(b)
        pr_n.append(t.shape[0])
    out = {"recall": {k: [] for k in ks}, "first_match": []}
    if not gt_t:
        return out
    first, _ = ops.sgg_match(torch.cat(gt_t), torch.cat(gt_b), gt_n, torch.cat(pr_t), torch.cat(pr_b), pr_n, iou_thres)
    first_host = first.cpu()
    off = 0
    for n in gt_n:
        fm = first_host[off:off + n]
        off += n
        out["first_match"].append(fm)
        for k in ks:
            out["recall"][k].append(float((fm < k).sum()) / float(n))                            # :249-252
    return out


def zeroshot_recall(first_match: Sequence[torch.Tensor], groundtruths, zeroshot_triplets: torch.Tensor,
                    ks: Sequence[int] = (20, 50, 100)) -> Dict[int, List[float]]:
    """SGZeroShotRecall (sgg_eval.py:277-309): recall restricted to the ground-truth triplets whose (subject class,
    object class, predicate) is in `zeroshot_triplets` [Z,3]; images without such a triplet contribute nothing.
    `first_match`: recall_at_k(...)['first_match'] for the same (non-empty) images, in order."""
    out = {k: [] for k in ks}
    zs = zeroshot_triplets.long().cpu()
    it = iter(first_match)
    for gt in groundtruths:
        rel = gt.get_field("relation_tuple").long().cpu()
        if rel.shape[0] == 0:
            continue
        fm = next(it).cpu()
        cls = gt.get_field("labels").long().cpu()
        trip = torch.stack((cls[rel[:, 0]], cls[rel[:, 1]], rel[:, 2]), 1)                       # :283-285
        is_zs = (trip[:, None, :] == zs[None, :, :]).all(-1).any(-1)                             # :287
        n = int(is_zs.sum())
        if n:
            for k in ks:
                out[k].append(float(((fm < k) & is_zs).sum()) / float(n))                        # :299-306
    return out


def pair_accuracy(predictions, groundtruths, ks: Sequence[int] = (20, 50, 100), iou_thres: float = 0.5,
                  predcls_like: bool = True) -> Dict[str, Dict[int, List[float]]]:
    """SGPairAccuracy (sgg_eval.py:338-366, PredCls / SGCls): matching restricted to the predictions whose (subject,
    object) pair is a ground-truth pair, ranks counted within that filtered list.  Returns {'hit': {k: [...]},
    'count': {k: [...]}} per image (ground-truth triplets matched within the first k filtered predictions; #gt)."""
    gt_t, gt_b, pr_t, pr_b, gt_n, pr_n = [], [], [], [], [], []
    for pred, gt in zip(predictions, groundtruths):
        rel_tuple = gt.get_field("relation_tuple").long()
        if rel_tuple.shape[0] == 0:
            continue
        gt_cls, gt_box = gt.get_field("labels").long(), gt.convert("xyxy").bbox
        t, b = triplets(rel_tuple[:, :2], rel_tuple[:, 2], gt_cls, gt_box)
        gt_t.append(t)
        gt_b.append(b)
        gt_n.append(t.shape[0])
        pairs = pred.get_field("rel_pair_idxs").long()
        keep = ((pairs[:, 0] * 1024 + pairs[:, 1])[:, None] == (rel_tuple[:, 0] * 1024 + rel_tuple[:, 1])[None, :]).any(1)  # :339-346
        scores = pred.get_field("pred_rel_scores")[keep]
        cls, box = (gt_cls, gt_box) if predcls_like else (pred.get_field("pred_labels").long(), pred.convert("xyxy").bbox)
        t, b = triplets(pairs[keep], 1 + scores[:, 1:].argmax(1), cls, box)
        pr_t.append(t)
        pr_b.append(b)
        pr_n.append(t.shape[0])
    out = {"hit": {k: [] for k in ks}, "count": {k: [] for k in ks}}
    if not gt_t:
        return out
    first, _ = ops.sgg_match(torch.cat(gt_t), torch.cat(gt_b), gt_n, torch.cat(pr_t), torch.cat(pr_b), pr_n, iou_thres)
    first_host, off = first.cpu(), 0
    for n in gt_n:
        fm = first_host[off:off + n]
        off += n
        for k in ks:
            out["hit"][k].append(float((fm < k).sum()))
            out["count"][k].append(float(n))
    return out
