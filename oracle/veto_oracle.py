"""ORACLE — test infrastructure only.  Never imported by veto_b200 (the product path).

A CPU (numpy, fp32) restatement of the reference's VETO relation-prediction hot path, written in
the reference's own formulation (pair tensors materialised, un-factored patch embedding, full last
layer) so that it is an independent check of the B200 kernels, which use a different factorisation.

Every function cites the reference lines it follows (paths relative to /root/reference).  The oracle
is pinned against outputs of the unmodified reference run under import shims
(tests/golden/make_golden.py writes tests/golden/*.npz; tests/test_oracle.py checks them); the
reference's own test-suite holds no golden vector for this path, so those fixtures are
the pin.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
from scipy.special import erf as _erf

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liboracle_roialign.so")
_lib = None

f32 = np.float32


def build_c(force: bool = False) -> str:
    """gcc-compile the C part of the oracle (ROIAlign)."""
    src = os.path.join(_HERE, "roialign_oracle.c")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", _LIB, src, "-lm"])
    return _LIB


def _clib():
    global _lib
    if _lib is None:
        build_c()
        _lib = ctypes.CDLL(_LIB)
        fp = ctypes.POINTER(ctypes.c_float)
        for name in ("oracle_roi_align_forward", "oracle_roi_align_backward"):
            fn = getattr(_lib, name)
            fn.restype = ctypes.c_int
            fn.argtypes = [fp, ctypes.c_int, ctypes.c_int, ctypes.c_int, fp, ctypes.c_int,
                           ctypes.c_float, ctypes.c_int, ctypes.c_int, ctypes.c_int, fp]
    return _lib


def _fp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


# --------------------------------------------------------------------------------------
# a1: candidate pair enumeration   (relation_head/sampling.py:31-52)
# --------------------------------------------------------------------------------------

def boxlist_iou(b1: np.ndarray, b2: np.ndarray) -> np.ndarray:
    """structures/boxlist_ops.py:54-87 (xyxy, +1 convention)."""
    a1 = (b1[:, 2] - b1[:, 0] + 1) * (b1[:, 3] - b1[:, 1] + 1)
    a2 = (b2[:, 2] - b2[:, 0] + 1) * (b2[:, 3] - b2[:, 1] + 1)
    lt = np.maximum(b1[:, None, :2], b2[None, :, :2])
    rb = np.minimum(b1[:, None, 2:], b2[None, :, 2:])
    wh = np.clip(rb - lt + 1, 0, None)
    inter = wh[..., 0] * wh[..., 1]
    return (inter / (a1[:, None] + a2[None, :] - inter)).astype(f32)


def prepare_test_pairs(n_boxes: Sequence[int], max_pairs: int = 2048,
                       scores: Optional[Sequence[np.ndarray]] = None,
                       boxes: Optional[Sequence[np.ndarray]] = None,
                       require_overlap: bool = False) -> List[np.ndarray]:
    """All ordered pairs (i,j), i!=j, row-major == torch.nonzero(ones-eye) (sampling.py:35-40).

    Over the cap, the reference keeps the top `max_pairs` by pred_scores[i]*pred_scores[j] with an
    unstable torch.sort (sampling.py:41-45).  Each pair ties with its mirror, so the order is only
    defined up to tie-breaking; this oracle (and the CUDA path) define the key as (score desc,
    row-major index asc) — what torch's CPU sort produces in practice."""
    out = []
    for k, n in enumerate(n_boxes):
        cand = np.ones((n, n), bool) & ~np.eye(n, dtype=bool)
        if require_overlap:
            cand &= boxlist_iou(boxes[k], boxes[k]) > 0
        idx = np.argwhere(cand).astype(np.int64).reshape(-1, 2)
        if len(idx) > max_pairs:
            s = scores[k].astype(f32)
            q = s[idx[:, 0]] * s[idx[:, 1]]
            sel = np.argsort(-q, kind="stable")[:max_pairs]
            idx = idx[sel]
        if len(idx) == 0:
            idx = np.zeros((1, 2), np.int64)        # sampling.py:47-51 placeholder
        out.append(idx)
    return out


# --------------------------------------------------------------------------------------
# a2/a3: ROI format, FPN level mapping, ROIAlign gather   (modeling/poolers.py)
# --------------------------------------------------------------------------------------

def rois_format(boxes: Sequence[np.ndarray]) -> np.ndarray:
    """poolers.py:96-107 — [N,5] fp32 (image index, x1, y1, x2, y2)."""
    rows = [np.concatenate([np.full((len(b), 1), i, f32), b.astype(f32)], 1) for i, b in enumerate(boxes)]
    return np.concatenate(rows, 0)


def level_map(boxes: Sequence[np.ndarray], k_min: int = 2, k_max: int = 5) -> np.ndarray:
    """poolers.py:32-43 with BoxList.area (+1, bounding_box.py:249-253); all fp32."""
    b = np.concatenate(list(boxes), 0).astype(f32)
    area = (b[:, 2] - b[:, 0] + f32(1)) * (b[:, 3] - b[:, 1] + f32(1))
    s = np.sqrt(area)
    lvl = np.floor(f32(4) + np.log2(s / f32(224) + f32(1e-6)))
    return (np.clip(lvl, k_min, k_max).astype(np.int64) - k_min)


def roi_align(inp: np.ndarray, rois: np.ndarray, scale: float, ph: int = 8, pw: int = 8, sr: int = 2) -> np.ndarray:
    """csrc/cpu/ROIAlign_cpu.cpp:114-219 through the C restatement."""
    inp = np.ascontiguousarray(inp, f32)
    rois = np.ascontiguousarray(rois, f32)
    B, C, H, W = inp.shape
    out = np.empty((len(rois), C, ph, pw), f32)
    if len(rois):
        rc = _clib().oracle_roi_align_forward(_fp(inp), C, H, W, _fp(rois), len(rois), scale, ph, pw, sr, _fp(out))
        assert rc == 0
    return out


def roi_align_backward(grad: np.ndarray, rois: np.ndarray, shape, scale: float, sr: int = 2) -> np.ndarray:
    """csrc/cuda/ROIAlign_cuda.cu:178-254 (serial accumulation instead of atomics)."""
    grad = np.ascontiguousarray(grad, f32)
    rois = np.ascontiguousarray(rois, f32)
    B, C, H, W = shape
    gi = np.zeros(shape, f32)
    rc = _clib().oracle_roi_align_backward(_fp(grad), C, H, W, _fp(rois), len(rois), scale,
                                           grad.shape[2], grad.shape[3], sr, _fp(gi))
    assert rc == 0
    return gi


def pooler_forward(feats: Sequence[np.ndarray], depth: np.ndarray, boxes: Sequence[np.ndarray],
                   scales=(0.25, 0.125, 0.0625, 0.03125), res: int = 8, sr: int = 2):
    """poolers.py:109-171 (depth + RGB branch) as called by VETOFeatureExtractor.forward
    (box_head/roi_box_feature_extractors.py:116-135): returns (x_2d, d_2d), both [N,256,8,8]."""
    rois = rois_format(boxes)
    lv = level_map(boxes)
    C = feats[0].shape[1]
    x2d = np.zeros((len(rois), C, res, res), f32)
    d2d = roi_align(depth, rois, scales[2], res, res, sr)          # always the level-2 pooler
    for l, (f, s) in enumerate(zip(feats, scales)):
        idx = np.nonzero(lv == l)[0]
        if len(idx):
            x2d[idx] = roi_align(f, rois[idx], s, res, res, sr)
    return x2d, d2d


# --------------------------------------------------------------------------------------
# a5-a9: predictor
# --------------------------------------------------------------------------------------

def _linear(x, w, b=None):
    y = x @ w.T
    if b is not None:
        y = y + b
    return y.astype(f32, copy=False)


def _layer_norm(x, w, b, eps=1e-5):
    mu = x.mean(-1, keepdims=True, dtype=np.float64)
    var = ((x - mu) ** 2).mean(-1, keepdims=True, dtype=np.float64)
    return (((x - mu) / np.sqrt(var + eps)) * w + b).astype(f32)


def _gelu(x):
    return (0.5 * x * (1.0 + _erf(x.astype(np.float64) / np.sqrt(2.0)))).astype(f32)


def _softmax(x, axis=-1):
    m = x.max(axis, keepdims=True)
    e = np.exp(x - m)
    return (e / e.sum(axis, keepdims=True)).astype(f32)


def center_xywh(boxes_xyxy: np.ndarray) -> np.ndarray:
    """BoxList.convert('xywh') (+1, bounding_box.py:72-75) then center_xywh (model_mpv2.py:342-345)."""
    b = boxes_xyxy.astype(f32)
    w = b[:, 2] - b[:, 0] + f32(1)
    h = b[:, 3] - b[:, 1] + f32(1)
    return np.stack([b[:, 0] + f32(0.5) * w, b[:, 1] + f32(0.5) * h, w, h], 1).astype(f32)


def box_embeddings(sd: Dict[str, np.ndarray], batch: Dict, mode: str, prefix: str = "",
                   meet: bool = False, obj_preds: Optional[np.ndarray] = None):
    """roi_relation_predictors.py:4081-4102 (eval): returns (obj_embed [N,200], pos_embed [N,128])."""
    P = prefix
    if mode == "predcls":
        labels = np.concatenate(batch["labels"])
        obj_embed = sd[P + "obj_embed.weight"][labels]
    elif meet:
        # Ensemble: hard lookup of predicted labels (roi_relation_predictors.py:3771-3784)
        obj_embed = sd[P + "obj_embed.weight"][obj_preds]
    else:
        logits = np.concatenate(batch["predict_logits"]).astype(f32)
        obj_embed = _softmax(logits, 1) @ sd[P + "obj_embed.weight"]
    cx = center_xywh(np.concatenate(batch["boxes"]))
    # BatchNorm1d(4) eval (running stats), Linear(4,128), ReLU; Dropout is identity in eval
    bn = (cx - sd[P + "pos_embed.0.running_mean"]) / np.sqrt(sd[P + "pos_embed.0.running_var"] + f32(1e-5))
    bn = (bn * sd[P + "pos_embed.0.weight"] + sd[P + "pos_embed.0.bias"]).astype(f32)
    pos = np.maximum(_linear(bn, sd[P + "pos_embed.1.weight"], sd[P + "pos_embed.1.bias"]), 0)
    return obj_embed.astype(f32), pos.astype(f32)


def global_pair_indices(pairs: Sequence[np.ndarray], n_boxes: Sequence[int]):
    """roi_relation_predictors.py:4104-4115 — per-image box offsets added to local indices."""
    subj, obj, off = [], [], 0
    for p, n in zip(pairs, n_boxes):
        subj.append(p[:, 0] + off)
        obj.append(p[:, 1] + off)
        off += n
    return np.concatenate(subj), np.concatenate(obj)


def patch_embed(sd, d, v, T="fusion_transformer.transformer."):
    """model_veto.py:99-115: 'b c (h p1) (w p2) -> b (h w) (p1 p2 c)', p=2; proj_d 2048->512 on the
    first argument (depth), proj_v 2048->64 on the second (RGB); cat -> 576."""
    def rearr(t):
        b, c, Hh, Ww = t.shape
        t = t.reshape(b, c, Hh // 2, 2, Ww // 2, 2)           # b c h p1 w p2
        t = t.transpose(0, 2, 4, 3, 5, 1)                      # b h w p1 p2 c
        return t.reshape(b, (Hh // 2) * (Ww // 2), 4 * c)
    pd = _linear(rearr(d), sd[T + "patch_embed.proj_d.weight"], sd[T + "patch_embed.proj_d.bias"])
    pv = _linear(rearr(v), sd[T + "patch_embed.proj_v.weight"], sd[T + "patch_embed.proj_v.bias"])
    return np.concatenate([pd, pv], 2)


def encoder(sd, x, T="fusion_transformer.transformer.", heads=6, layers=6, return_all=False):
    """model_veto.py:15-26 loop; PreNorm :125-132; Attention :67-96; FeedForward :134-146."""
    b, n, D = x.shape
    dh = D // heads
    scale = f32(dh ** -0.5)
    for i in range(layers):
        L = f"{T}layers.{i}."
        xn = _layer_norm(x, sd[L + "0.norm.weight"], sd[L + "0.norm.bias"])
        qkv = _linear(xn, sd[L + "0.fn.to_qkv.weight"])
        q, k, v = [t.reshape(b, n, heads, dh).transpose(0, 2, 1, 3) for t in np.split(qkv, 3, -1)]
        dots = (q @ k.transpose(0, 1, 3, 2)) * scale
        attn = _softmax(dots, -1)
        o = (attn @ v).transpose(0, 2, 1, 3).reshape(b, n, D)
        x = _linear(o, sd[L + "0.fn.to_out.0.weight"], sd[L + "0.fn.to_out.0.bias"]) + x
        xn = _layer_norm(x, sd[L + "1.norm.weight"], sd[L + "1.norm.bias"])
        h = _gelu(_linear(xn, sd[L + "1.fn.net.0.weight"], sd[L + "1.fn.net.0.bias"]))
        x = _linear(h, sd[L + "1.fn.net.3.weight"], sd[L + "1.fn.net.3.bias"]) + x
    return x if return_all else x[:, 0]


def relation_features(sd, batch, pairs, x2d, d2d, mode, prefix="", meet=False, obj_preds=None,
                      chunk: int = 512) -> np.ndarray:
    """roi_relation_predictors.py:4104-4124 + model_veto.py:52-64: returns the CLS vector [R,576]."""
    P = prefix
    T = P + "fusion_transformer.transformer."
    obj_embed, pos = box_embeddings(sd, batch, mode, P, meet, obj_preds)
    s, o = global_pair_indices(pairs, batch["n_boxes"])
    out = np.empty((len(s), 576), f32)
    for a in range(0, len(s), chunk):
        si, oi = s[a:a + chunk], o[a:a + chunk]
        loc = np.maximum(_linear(np.concatenate([pos[si], pos[oi]], 1),
                                 sd[P + "location_projection.0.weight"], sd[P + "location_projection.0.bias"]), 0)
        cls = np.maximum(_linear(np.concatenate([obj_embed[si], obj_embed[oi]], 1),
                                 sd[P + "class_projection.0.weight"], sd[P + "class_projection.0.bias"]), 0)
        rel_visual = np.concatenate([x2d[si], x2d[oi]], 1)       # [r,512,8,8]
        rel_depth = np.concatenate([d2d[si], d2d[oi]], 1)
        patches = patch_embed(sd, rel_depth, rel_visual, T)      # fusion_transformer(rel_depth, rel_visual, ...)
        r = len(si)
        tok = np.concatenate([np.broadcast_to(sd[T + "cls_token"], (r, 1, 576)), patches,
                              loc[:, None], cls[:, None]], 1)
        tok = (tok + sd[T + "pos_embedding"]).astype(f32)        # same vector on every token (:62)
        out[a:a + chunk] = encoder(sd, tok, T)
    return out


def tokens_only(sd, batch, pairs, x2d, d2d, mode, prefix=""):
    """The [R,19,576] token tensor before the encoder (for stage-level parity tests)."""
    P = prefix
    T = P + "fusion_transformer.transformer."
    obj_embed, pos = box_embeddings(sd, batch, mode, P)
    s, o = global_pair_indices(pairs, batch["n_boxes"])
    loc = np.maximum(_linear(np.concatenate([pos[s], pos[o]], 1),
                             sd[P + "location_projection.0.weight"], sd[P + "location_projection.0.bias"]), 0)
    cls = np.maximum(_linear(np.concatenate([obj_embed[s], obj_embed[o]], 1),
                             sd[P + "class_projection.0.weight"], sd[P + "class_projection.0.bias"]), 0)
    patches = patch_embed(sd, np.concatenate([d2d[s], d2d[o]], 1), np.concatenate([x2d[s], x2d[o]], 1), T)
    tok = np.concatenate([np.broadcast_to(sd[T + "cls_token"], (len(s), 1, 576)), patches, loc[:, None], cls[:, None]], 1)
    return (tok + sd[T + "pos_embedding"]).astype(f32)


def predictor_forward(sd, batch, pairs, x2d, d2d, mode="predcls") -> np.ndarray:
    """VETOPredictor.forward eval (roi_relation_predictors.py:4074-4139): rel logits [R,C_rel]."""
    feat = relation_features(sd, batch, pairs, x2d, d2d, mode)
    return _linear(feat, sd["rel_out.weight"], sd["rel_out.bias"])


def obj_dists_onehot(batch, mode, num_obj):
    """F.one_hot(obj_labels) floats (:4088,4092)."""
    lab = np.concatenate(batch["labels"] if mode == "predcls" else batch["pred_labels"])
    return np.eye(num_obj, dtype=f32)[lab]


def incre_idx_list(group_sizes: Sequence[int], num_rel: int) -> List[int]:
    """SHA_GCL_extra/extra_function_utils.py:39-52: predicate id -> 1-based group id (0 for bg)."""
    out = [0] * num_rel
    c = 1
    for g, n in enumerate(group_sizes):
        for _ in range(n):
            out[c] = g + 1
            c += 1
    return out


def softmax_rows(x: np.ndarray) -> np.ndarray:
    e = np.exp(x - x.max(1, keepdims=True), dtype=np.float32)
    return e / e.sum(1, keepdims=True, dtype=np.float32)


def meet_forward(sd, batch, pairs, x2d, d2d, group_sizes, mode="predcls", nms_thresh: float = 0.5,
                 nms_scores: Optional[np.ndarray] = None) -> Dict[str, np.ndarray]:
    """VETOPredictor_MEET.forward eval (:3909-3995) -> Ensemble.forward (:3752-3853),
    EXPERT_GROUP False: {'group_k': [R, n_k+2]} un-split.  sgdet test with boxes_per_cls: obj_preds from nms_per_cls
    over softmax(one_hot(pred_labels)) (:3776-3781; `nms_scores` overrides that softmax, see tests)."""
    obj_preds = None
    if mode == "sgdet" and "boxes_per_cls" in batch:
        labels = np.concatenate(batch["pred_labels"])
        if nms_scores is None:
            nms_scores = softmax_rows(np.eye(batch["num_obj"], dtype=np.float32)[labels])
        obj_preds = nms_per_cls(nms_scores, batch["boxes_per_cls"], batch["n_boxes"], nms_thresh)
    elif mode != "predcls":
        # non-sgdet-test branch: obj_dists one-hot -> argmax over [1:] (+1) (:3783)
        obj_preds = np.concatenate(batch["pred_labels"])
    feat = relation_features(sd, batch, pairs, x2d, d2d, mode, prefix="model.", meet=True, obj_preds=obj_preds)
    if "model.rel_out_group.0.0.weight" in sd:      # EXPERT_GROUP: 'group_%d%d' % (k, expert + 1) (:3834-3840)
        experts = 1 + max(int(key.split(".")[2]) for key in sd if key.startswith("model.rel_out_group."))
        return {"group_%d%d" % (k, j + 1): _linear(feat, sd[f"model.rel_out_group.{j}.{k}.weight"],
                                                   sd[f"model.rel_out_group.{j}.{k}.bias"])
                for j in range(experts) for k in range(len(group_sizes))}
    return {f"group_{k}": _linear(feat, sd[f"model.rel_out.{k}.weight"], sd[f"model.rel_out.{k}.bias"])
            for k in range(len(group_sizes))}


def cross_entropy(logits: np.ndarray, labels: np.ndarray, weight: Optional[np.ndarray] = None) -> float:
    """nn.CrossEntropyLoss(weight) mean reduction (:4070,4134-4135)."""
    z = logits.astype(np.float64)
    z = z - z.max(1, keepdims=True)
    lp = z - np.log(np.exp(z).sum(1, keepdims=True))
    w = np.ones(logits.shape[1]) if weight is None else weight.astype(np.float64)
    wl = w[labels]
    return float(-(wl * lp[np.arange(len(labels)), labels]).sum() / wl.sum())


# --------------------------------------------------------------------------------------
# a12: PostProcessor, vanilla branch   (relation_head/inference.py:398-453)
# --------------------------------------------------------------------------------------

def postprocess(rel_logits: Sequence[np.ndarray], obj_logits: Sequence[np.ndarray], pairs: Sequence[np.ndarray]):
    """use_gt_box branch: obj scores = max softmax[:,1:]; triple = rel*s0*s1; sort desc.
    torch.sort is unstable; ties are broken here by original index asc (stable), and callers
    compare rankings tie-aware."""
    res = []
    for rl, ol, pr in zip(rel_logits, obj_logits, pairs):
        op = _softmax(ol.astype(f32), -1)
        op[:, 0] = 0
        obj_scores = op[:, 1:].max(1)
        obj_pred = op[:, 1:].argmax(1) + 1
        rp = _softmax(rl.astype(f32), -1)
        rel_scores = rp[:, 1:].max(1)
        rel_class = rp[:, 1:].argmax(1) + 1
        triple = (rel_scores * obj_scores[pr[:, 0]] * obj_scores[pr[:, 1]]).astype(f32)
        order = np.argsort(-triple, kind="stable")
        res.append(dict(rel_pair_idxs=pr[order], pred_rel_scores=rp[order], pred_rel_labels=rel_class[order],
                        triple_scores=triple[order], pred_labels=obj_pred, pred_scores=obj_scores))
    return res


def nms_overlaps(boxes: np.ndarray) -> np.ndarray:
    """relation_head/utils_relation.py:56-79: per-class IoU [n, n, C] of boxes [n, C, 4] (xyxy, +1 convention), fp32 in
    the reference's operation order."""
    b = boxes.astype(np.float32)
    max_xy = np.minimum(b[:, None, :, 2:], b[None, :, :, 2:])
    min_xy = np.maximum(b[:, None, :, :2], b[None, :, :, :2])
    inter = np.clip(max_xy - min_xy + np.float32(1.0), 0, None)
    inters = inter[..., 0] * inter[..., 1]
    areas = (b[..., 2] - b[..., 0] + np.float32(1.0)) * (b[..., 3] - b[..., 1] + np.float32(1.0))
    union = -inters + areas[None] + areas[:, None]
    return inters / union


def nms_per_cls(scores: np.ndarray, boxes_per_cls: Sequence[np.ndarray], n_boxes: Sequence[int], thresh: float) -> np.ndarray:
    """Ensemble.nms_per_cls (roi_relation_predictors.py:3855-3874).  `scores` [N, C] = softmax of the object
    distribution (the reference computes F.softmax(obj_dists[i], -1) per image and takes it to numpy)."""
    out, off = [], 0
    for i, n in enumerate(n_boxes):
        is_overlap = nms_overlaps(boxes_per_cls[i]) >= np.float32(thresh)
        sampled = scores[off:off + n].astype(np.float32).copy()
        off += n
        sampled[:, 0] = -1
        label = np.zeros(n, np.int64)
        for _ in range(n):
            box_ind, cls_ind = np.unravel_index(sampled.argmax(), sampled.shape)
            label[int(box_ind)] = int(cls_ind)
            sampled[is_overlap[box_ind, :, cls_ind], cls_ind] = 0.0
            sampled[box_ind] = -1.0
        out.append(label)
    return np.concatenate(out) if out else np.zeros(0, np.int64)


def obj_prediction_nms(scores: np.ndarray, boxes_per_cls: np.ndarray, thresh: float) -> np.ndarray:
    """relation_head/utils_relation.py:94-128 for ONE image: `scores` = softmax(pred_logits) [n, C]."""
    n = scores.shape[0]
    is_overlap = nms_overlaps(boxes_per_cls) >= np.float32(thresh)
    prob = scores.astype(np.float32).copy()
    prob[:, 0] = 0
    label = np.zeros(n, np.int64)
    for _ in range(n):
        box_ind, cls_ind = np.unravel_index(prob.argmax(), prob.shape)
        if not label[int(box_ind)] > 0:
            label[int(box_ind)] = int(cls_ind)
        prob[is_overlap[box_ind, :, cls_ind], cls_ind] = 0.0
        prob[box_ind] = -1.0
    return label


def postprocess_sgdet(rel_logits, obj_logits, pairs, boxes_per_cls, thresh: float, obj_scores_softmax=None):
    """PostProcessor.forward, vanilla branch with use_gt_box False (relation_head/inference.py:398-453): per image
    dict(obj_pred, obj_scores, boxes, pairs, probs, labels, triple), sorted by triple score (stable)."""
    res = []
    for i, (rl, ol, pr, bpc) in enumerate(zip(rel_logits, obj_logits, pairs, boxes_per_cls)):
        op = softmax_rows(ol.astype(np.float32)) if obj_scores_softmax is None else obj_scores_softmax[i].copy()
        op[:, 0] = 0
        pred = obj_prediction_nms(op, bpc, thresh)
        sc = op[np.arange(len(pred)), pred]
        rp = softmax_rows(rl.astype(np.float32))
        rs, rc = rp[:, 1:].max(1), rp[:, 1:].argmax(1) + 1
        triple = rs * sc[pr[:, 0]] * sc[pr[:, 1]]
        order = np.argsort(-triple, kind="stable")
        res.append(dict(obj_pred=pred, obj_scores=sc, boxes=bpc[np.arange(len(pred)), pred], pairs=pr[order],
                        probs=rp[order], labels=rc[order], triple=triple[order]))
    return res


def postprocess_meet(group_logits: Dict[str, np.ndarray], obj_logits: np.ndarray, pairs: np.ndarray, incre_idx: Sequence[int]):
    """PostProcessor.forward, 'ensemble' branch (relation_head/inference.py:284-397) for ONE image in PredCls / SGCls
    mode (use_gt_box): returns dict(pairs [G*R,2], probs [G*R,num_rel], labels [G*R] head-local, triple [G*R]) ranked by
    triple score (stable)."""
    op = softmax_rows(obj_logits.astype(np.float32))
    op[:, 0] = 0
    sc = op[:, 1:].max(1)
    num_rel = len(incre_idx)
    triple, prs, labs, probs = [], [], [], []
    for k in range(len(group_logits)):
        p = softmax_rows(group_logits["group_%d" % k].astype(np.float32))[:, :-1]
        rs, rc = p[:, 1:].max(1), p[:, 1:].argmax(1) + 1
        cols = [0] + [i for i, g in enumerate(incre_idx) if g == k + 1]
        full = np.zeros((len(p), num_rel), np.float32)
        full[:, cols] = p
        triple.append(rs * sc[pairs[:, 0]] * sc[pairs[:, 1]])
        prs.append(pairs)
        labs.append(rc)
        probs.append(full)
    triple = np.concatenate(triple)
    order = np.argsort(-triple, kind="stable")
    return dict(pairs=np.concatenate(prs)[order], probs=np.concatenate(probs)[order], labels=np.concatenate(labs)[order],
                triple=triple[order])


def gtbox_relsample_candidates(rel: np.ndarray):
    """The deterministic part of RelationSampling.gtbox_relsample (relation_head/sampling.py:54-107) for one image:
    (foreground pairs in nonzero() order [F,2], their labels [F], background candidates [G,2] = every other ordered
    pair i != j, symmetric binary matrix [n,n]).  What the reference then draws at random: a num_pos subset of the
    foreground rows if there are more (:91-94) and a random permutation of the background rows cut to
    batch_size - num_fg (:97-99)."""
    n = rel.shape[0]
    fg = np.argwhere(rel > 0).astype(np.int64).reshape(-1, 2)
    labels = rel[fg[:, 0], fg[:, 1]].astype(np.int64)
    binary = np.zeros((n, n), np.int64)
    binary[fg[:, 0], fg[:, 1]] = 1
    binary[fg[:, 1], fg[:, 0]] = 1
    poss = np.ones((n, n), np.int64) - np.eye(n, dtype=np.int64)
    poss[fg[:, 0], fg[:, 1]] = 0
    bg = np.argwhere(poss > 0).astype(np.int64).reshape(-1, 2)
    return fg, labels, bg, binary


def compute_pred_matches(gt_triplets, pred_triplets, gt_boxes, pred_boxes, iou_thres: float):
    """_compute_pred_matches (data/datasets/evaluation/vg/sgg_eval.py:77-117), non-phrdet: pred_to_gt as a list of
    lists (ground-truth indices matched by each prediction, ascending)."""
    keeps = (gt_triplets[..., None] == pred_triplets.T[None, ...]).all(1)     # intersect_2d, utils/miscellaneous.py:47-61
    pred_to_gt = [[] for _ in range(pred_boxes.shape[0])]
    for g in np.where(keeps.any(1))[0]:
        idx = np.where(keeps[g])[0]
        sub = boxlist_iou(gt_boxes[g:g + 1, :4].astype(f32), pred_boxes[idx, :4].astype(f32))[0]
        obj = boxlist_iou(gt_boxes[g:g + 1, 4:].astype(f32), pred_boxes[idx, 4:].astype(f32))[0]
        for i in idx[(sub >= iou_thres) & (obj >= iou_thres)]:
            pred_to_gt[int(i)].append(int(g))
    return pred_to_gt


def recall_at_k(pred_to_gt, n_gt: int, ks=(20, 50, 100)):
    """SGRecall.calculate_recall (:155-160): recall@k = |union of pred_to_gt[:k]| / n_gt."""
    out = {}
    for k in ks:
        match = set()
        for m in pred_to_gt[:k]:
            match |= set(m)
        out[k] = len(match) / float(n_gt)
    return out


def detect_relsample_candidates(prp_boxes, prp_labels, prp_scores, tgt_boxes, tgt_labels, relation, fg_thres: float = 0.5,
                                require_overlap: bool = False):
    """The deterministic part of RelationSampling.detect_relsample / motif_rel_fg_bg_sampling
    (relation_head/sampling.py:109-309) for one image: dict(locating [P], binary [P,P], gt = [(head, tail, label,
    candidate pairs [(a,b)...] in head-major order)] in nonzero() order, bg = background pairs [(a,b)...] sorted by
    pred_scores[a]*pred_scores[b] descending (stable), the pool the reference cuts to 2*num_neg and then draws from).
    What the reference draws at random: <= num_sample_per_gt_rel candidates per ground-truth relation with probability
    ~ iou_head*iou_tail (:256-261), a num_pos subset of all foreground rows (:271-273), num_neg of the pool (:290-291)."""
    P = len(prp_boxes)
    ious = boxlist_iou(tgt_boxes.astype(f32), prp_boxes.astype(f32))                   # [T, P]
    is_match = (tgt_labels[:, None] == prp_labels[None]) & (ious > f32(fg_thres))
    locating = (ious > f32(fg_thres)).any(0).astype(f32)
    if require_overlap:
        self_iou = boxlist_iou(prp_boxes.astype(f32), prp_boxes.astype(f32))
        poss = (self_iou > 0) & (self_iou < 1)
    else:
        poss = ~np.eye(P, dtype=bool)
    poss = poss.copy()
    poss[prp_labels == 0] = False
    poss[:, prp_labels == 0] = False
    binary = np.zeros((P, P), np.int64)
    gt = []
    for h, t in np.argwhere(relation != 0):
        heads, tails = np.nonzero(is_match[h])[0], np.nonzero(is_match[t])[0]
        if len(heads) and len(tails):
            binary[np.ix_(heads, tails)] = 1
            binary[np.ix_(tails, heads)] = 1
        pairs = [(int(a), int(b)) for a in heads for b in tails if a != b]
        for a, b in pairs:
            poss[a, b] = False
        gt.append((int(h), int(t), int(relation[h, t]), pairs))
    bg = np.argwhere(poss)
    q = prp_scores.astype(f32)[bg[:, 0]] * prp_scores.astype(f32)[bg[:, 1]] if len(bg) else np.zeros(0, f32)
    bg = bg[np.argsort(-q, kind="stable")]
    return dict(locating=locating, binary=binary, gt=gt, bg=[(int(a), int(b)) for a, b in bg], ious=ious)


def postprocess_meet_vote(group_logits: Dict[str, np.ndarray], obj_logits: np.ndarray, pairs: np.ndarray, incre_idx: Sequence[int],
                          voting: str = "C"):
    """PostProcessor.forward, EXPERT_GROUP branch (relation_head/inference.py:93-283) for ONE image in PredCls / SGCls
    mode: three experts per group ('group_%d%d' % (group, expert 1..3)); voting 'C' = at least two experts agree on the
    class, 'U' = all three.  Returns dict(pairs, probs [.,num_rel], labels (head-local), triple) of the survivors, ranked
    by score (stable).  Follows the reference's arithmetic, including mean(p1, p1) for the expert pair (1,2) (:191-193)."""
    op = softmax_rows(obj_logits.astype(np.float32))
    op[:, 0] = 0
    sc = op[:, 1:].max(1)
    so = sc[pairs[:, 0]] * sc[pairs[:, 1]]
    num_rel = len(incre_idx)
    n_groups = len(group_logits) // 3
    T, PR, LB, PB = [], [], [], []
    for j in range(n_groups):
        p = [softmax_rows(group_logits["group_%d%d" % (j, e)].astype(np.float32))[:, :-1] for e in (1, 2, 3)]
        c = [x[:, 1:].argmax(1) + 1 for x in p]
        t = [x[:, 1:].max(1) * so for x in p]
        agree = [c[0] == c[1], c[1] == c[2], c[0] == c[2]]
        if voting == "U":
            keep = agree[0] & agree[1] & agree[2]
            triple = np.mean(np.stack(t, 1), 1)
            prob = np.mean(np.stack(p, 1), 1)
            cls = c[2]
        else:
            ab = np.stack(agree, 1)
            count = ab.sum(1)
            tavg = np.stack([(t[0] + t[1]) / 2, (t[1] + t[2]) / 2, (t[0] + t[2]) / 2], 1)
            pavg = np.stack([(p[0] + p[1]) / 2, (p[1] + p[1]) / 2, (p[0] + p[2]) / 2], 1)
            with np.errstate(invalid="ignore", divide="ignore"):
                triple = np.where(ab, tavg, 0).sum(1) / count
                prob = np.where(ab[:, :, None], pavg, 0).sum(1) / count[:, None]
            triple = np.nan_to_num(triple, nan=0.0)
            cls = np.zeros_like(c[0])
            for cc, a in zip(c, agree):
                cls[a] = cc[a]
            keep = ab.any(1)
        cols = [0] + [i for i, g in enumerate(incre_idx) if g == j + 1]
        full = np.zeros((len(pairs), num_rel), np.float32)
        full[:, cols] = prob
        T.append(triple[keep].astype(np.float32))
        PR.append(pairs[keep])
        LB.append(cls[keep])
        PB.append(full[keep])
    triple = np.concatenate(T)
    order = np.argsort(-triple, kind="stable")
    return dict(pairs=np.concatenate(PR)[order], probs=np.concatenate(PB)[order], labels=np.concatenate(LB)[order],
                triple=triple[order])
