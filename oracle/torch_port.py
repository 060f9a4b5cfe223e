"""ORACLE — test infrastructure / CPU baseline only.  Never imported by veto_b200 (the product path).

A PyTorch-CPU restatement of the reference's VETO relation-head path in the reference's OWN formulation and with the
same library calls it makes (torch.nonzero, torchvision.ops.roi_align(aligned=False) — bit-identical to the
reference's ROIAlign_cpu.cpp, F.linear / layer_norm / softmax / gelu on materialised pair tensors), so that timing it
on the host cores times the kernels the reference itself would run on CPU (MKL/oneDNN, all threads).  It exists
because the reference package cannot travel to the GPU box; it is pinned against the golden fixtures the unmodified
reference produced (tests/test_oracle.py::test_torch_port_matches_reference).

Reference lines (paths relative to /root/reference/pysgg/modeling):
  pairs     roi_heads/relation_head/sampling.py:31-52
  pooler    poolers.py:32-43, 96-171
  predictor roi_heads/relation_head/roi_relation_predictors.py:4081-4139
  encoder   roi_heads/relation_head/model_veto.py:15-26, 52-64, 67-96, 99-115, 125-146
  post      roi_heads/relation_head/inference.py:398-453
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np
import torch
import torch.nn.functional as F


def prepare_test_pairs(n_boxes: Sequence[int], max_pairs: int, scores=None) -> List[torch.Tensor]:
    out = []
    for k, n in enumerate(n_boxes):
        cand = torch.ones((n, n)) - torch.eye(n)
        idxs = torch.nonzero(cand).view(-1, 2)
        if len(idxs) > max_pairs:
            q = scores[k][idxs[:, 0]] * scores[k][idxs[:, 1]]
            idxs = idxs[torch.sort(q, descending=True, stable=True)[1][:max_pairs]]
        out.append(idxs if len(idxs) else torch.zeros((1, 2), dtype=torch.int64))
    return out


def pooler_forward(feats: Sequence[torch.Tensor], depth: torch.Tensor, boxes: Sequence[torch.Tensor],
                   scales=(0.25, 0.125, 0.0625, 0.03125), res: int = 8, sr: int = 2):
    import torchvision
    rois = torch.cat([torch.cat([torch.full((len(b), 1), float(i)), b], 1) for i, b in enumerate(boxes)], 0)
    allb = torch.cat(list(boxes), 0)
    s = torch.sqrt((allb[:, 2] - allb[:, 0] + 1) * (allb[:, 3] - allb[:, 1] + 1))
    lv = torch.clamp(torch.floor(4 + torch.log2(s / 224 + 1e-6)), min=2, max=5).to(torch.int64) - 2
    x2d = torch.zeros((len(rois), feats[0].shape[1], res, res))
    d2d = torchvision.ops.roi_align(depth, rois, (res, res), scales[2], sr, aligned=False)
    for l, (f, sc) in enumerate(zip(feats, scales)):
        idx = torch.nonzero(lv == l).squeeze(1)
        if len(idx):
            x2d[idx] = torchvision.ops.roi_align(f, rois[idx], (res, res), sc, sr, aligned=False)
    return x2d, d2d


def predictor_forward(sd: Dict[str, torch.Tensor], boxes: Sequence[torch.Tensor], pairs: Sequence[torch.Tensor],
                      x2d: torch.Tensor, d2d: torch.Tensor, mode: str, labels=None, predict_logits=None,
                      heads: int = 6, layers: int = 6, train: bool = False, drop=None, bn_out=None) -> torch.Tensor:
    """train=True: BatchNorm1d uses the batch statistics (and writes the updated running statistics into `bn_out`), and
    `drop` = {"pos": [N,128], "emb": [R,19,576], "attn": [layers x [R,19,576]]} holds explicit keep-scale factors
    (mask / (1 - p)) applied where the reference's nn.Dropout layers sit (roi_relation_predictors.py:4046,
    model_veto.py:63,83-85); None = no dropout."""
    T = "fusion_transformer.transformer."
    drop = drop or {}
    if mode == "predcls":
        obj_embed = sd["obj_embed.weight"][torch.cat(list(labels))]
    else:
        obj_embed = F.softmax(torch.cat(list(predict_logits)), 1) @ sd["obj_embed.weight"]
    b = torch.cat(list(boxes), 0)
    w, h = b[:, 2] - b[:, 0] + 1, b[:, 3] - b[:, 1] + 1
    cx = torch.stack([b[:, 0] + 0.5 * w, b[:, 1] + 0.5 * h, w, h], 1)
    if train:
        rm, rv = sd["pos_embed.0.running_mean"].clone(), sd["pos_embed.0.running_var"].clone()
        bn = F.batch_norm(cx, rm, rv, sd["pos_embed.0.weight"], sd["pos_embed.0.bias"], True, 0.001, 1e-5)
        if bn_out is not None:
            bn_out["running_mean"], bn_out["running_var"] = rm, rv
    else:
        bn = F.batch_norm(cx, sd["pos_embed.0.running_mean"], sd["pos_embed.0.running_var"], sd["pos_embed.0.weight"],
                          sd["pos_embed.0.bias"], False, 0.0, 1e-5)
    pos = F.relu(F.linear(bn, sd["pos_embed.1.weight"], sd["pos_embed.1.bias"]))
    if drop.get("pos") is not None:
        pos = pos * drop["pos"]
    subj, obj, off = [], [], 0
    for p, bx in zip(pairs, boxes):
        subj.append(p[:, 0] + off)
        obj.append(p[:, 1] + off)
        off += len(bx)
    s, o = torch.cat(subj), torch.cat(obj)
    loc = F.relu(F.linear(torch.cat((pos[s], pos[o]), 1), sd["location_projection.0.weight"], sd["location_projection.0.bias"]))
    cls = F.relu(F.linear(torch.cat((obj_embed[s], obj_embed[o]), 1), sd["class_projection.0.weight"],
                          sd["class_projection.0.bias"]))
    rel_visual = torch.cat((x2d[s], x2d[o]), 1)
    rel_depth = torch.cat((d2d[s], d2d[o]), 1)

    def rearr(t):  # 'b c (h p1) (w p2) -> b (h w) (p1 p2 c)', p = 2
        bb, c, H, W = t.shape
        return t.reshape(bb, c, H // 2, 2, W // 2, 2).permute(0, 2, 4, 3, 5, 1).reshape(bb, (H // 2) * (W // 2), 4 * c)

    pd = F.linear(rearr(rel_depth), sd[T + "patch_embed.proj_d.weight"], sd[T + "patch_embed.proj_d.bias"])
    pv = F.linear(rearr(rel_visual), sd[T + "patch_embed.proj_v.weight"], sd[T + "patch_embed.proj_v.bias"])
    x = torch.cat((pd, pv), 2)
    r = x.shape[0]
    x = torch.cat((sd[T + "cls_token"].expand(r, -1, -1), x, loc.unsqueeze(1), cls.unsqueeze(1)), 1) + sd[T + "pos_embedding"]
    if drop.get("emb") is not None:
        x = x * drop["emb"]
    D = x.shape[-1]
    dh = D // heads
    for i in range(layers):
        Lk = f"{T}layers.{i}."
        xn = F.layer_norm(x, (D,), sd[Lk + "0.norm.weight"], sd[Lk + "0.norm.bias"], 1e-5)
        q, k, v = [t.reshape(r, -1, heads, dh).permute(0, 2, 1, 3) for t in F.linear(xn, sd[Lk + "0.fn.to_qkv.weight"]).chunk(3, -1)]
        attn = torch.softmax(torch.einsum("bhid,bhjd->bhij", q, k) * dh ** -0.5, -1)
        out = torch.einsum("bhij,bhjd->bhid", attn, v).permute(0, 2, 1, 3).reshape(r, -1, D)
        proj = F.linear(out, sd[Lk + "0.fn.to_out.0.weight"], sd[Lk + "0.fn.to_out.0.bias"])
        if drop.get("attn") is not None:
            proj = proj * drop["attn"][i]
        x = proj + x
        xn = F.layer_norm(x, (D,), sd[Lk + "1.norm.weight"], sd[Lk + "1.norm.bias"], 1e-5)
        hdn = F.gelu(F.linear(xn, sd[Lk + "1.fn.net.0.weight"], sd[Lk + "1.fn.net.0.bias"]))
        x = F.linear(hdn, sd[Lk + "1.fn.net.3.weight"], sd[Lk + "1.fn.net.3.bias"]) + x
    return F.linear(x[:, 0], sd["rel_out.weight"], sd["rel_out.bias"])


def to_torch(sd_np: Dict[str, np.ndarray]) -> Dict[str, torch.Tensor]:
    return {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd_np.items()}


TRAINED_KEYS_EXCLUDED = ("obj_embed2.weight", "bbox_embed.", "running_", "num_batches_tracked", "criterion_loss")


def train_step(sd: Dict[str, torch.Tensor], boxes, pairs, rel_labels, x2d, d2d, mode: str, labels=None,
               predict_logits=None, class_weight=None, drop=None, layers: int = 6):
    """rel_loss of VETOPredictor.forward in train() mode (roi_relation_predictors.py:4131-4136) and, by autograd over
    this restatement, its gradients: returns (loss, {state_dict key: grad}, grad wrt d2d, grad wrt x2d, bn_out)."""
    leaves = {}
    for k, v in sd.items():
        if v.dtype.is_floating_point and not any(t in k for t in TRAINED_KEYS_EXCLUDED):
            leaves[k] = v.detach().clone().requires_grad_(True)
    sd2 = dict(sd)
    sd2.update(leaves)
    x2d = x2d.detach().clone().requires_grad_(True)
    d2d = d2d.detach().clone().requires_grad_(True)
    bn_out = {}
    logits = predictor_forward(sd2, boxes, pairs, x2d, d2d, mode, labels=labels, predict_logits=predict_logits,
                               layers=layers, train=True, drop=drop, bn_out=bn_out)
    loss = F.cross_entropy(logits, torch.cat(list(rel_labels)).long(), weight=class_weight)
    loss.backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in leaves.items()}
    return loss.detach(), grads, d2d.grad, x2d.grad, bn_out, logits.detach()


def train_step_meet(sd: Dict[str, torch.Tensor], boxes, pairs, head_labels, x2d, d2d, obj_preds, drop=None,
                    layers: int = 6):
    """VETOPredictor_MEET / Ensemble in train() mode (roi_relation_predictors.py:3806-3848, EXPERT_GROUP False): the
    trunk with hard class embeddings of `obj_preds`, one Linear head per group, and per head k the plain mean CE over
    the rows with head_labels[k] >= 0 (the pairs the group sampling chose, already relabelled group-locally); gradients
    of the SUM of the head losses by autograd.  `sd` holds the MEET state_dict keys ('model.' prefix).
    Returns ([loss_k], {state_dict key: grad}, grad wrt d2d, grad wrt x2d, bn_out)."""
    leaves = {}
    for k, v in sd.items():
        if v.dtype.is_floating_point and not any(t in k for t in TRAINED_KEYS_EXCLUDED):
            leaves[k] = v.detach().clone().requires_grad_(True)
    flat = {k[len("model."):]: leaves.get(k, v) for k, v in sd.items() if k.startswith("model.")}
    n_heads = len(head_labels)
    flat["rel_out.weight"] = torch.cat([flat[f"rel_out.{k}.weight"] for k in range(n_heads)], 0)
    flat["rel_out.bias"] = torch.cat([flat[f"rel_out.{k}.bias"] for k in range(n_heads)], 0)
    x2d = x2d.detach().clone().requires_grad_(True)
    d2d = d2d.detach().clone().requires_grad_(True)
    bn_out = {}
    logits = predictor_forward(flat, boxes, pairs, x2d, d2d, "predcls", labels=[obj_preds], layers=layers, train=True,
                               drop=drop, bn_out=bn_out)
    losses, col = [], 0
    for k in range(n_heads):
        n = flat[f"rel_out.{k}.weight"].shape[0]
        y = torch.as_tensor(head_labels[k]).long()
        rows = torch.nonzero(y >= 0)[:, 0]
        losses.append(F.cross_entropy(logits[rows, col:col + n], y[rows]))
        col += n
    torch.stack(losses).sum().backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in leaves.items()}
    return [l.detach() for l in losses], grads, d2d.grad, x2d.grad, bn_out
