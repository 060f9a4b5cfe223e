"""Functional wrappers over the C ABI (include/veto_b200.h).  Inputs and outputs are CUDA tensors; the
work is done by libveto_b200.so on the current stream.  No CPU fallback."""
from __future__ import annotations

import ctypes
import itertools
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import lib as L

T_DIM, N_TOKENS = 576, 19


def _i32(seq: Sequence[int]):
    return (ctypes.c_int32 * len(seq))(*[int(v) for v in seq])


def offsets_tensor(counts: Sequence[int], device) -> torch.Tensor:
    """int32 prefix sums [len+1] on `device` (pinned staging, asynchronous copy)."""
    off = [0]
    for c in counts:
        off.append(off[-1] + int(c))
    t = torch.tensor(off, dtype=torch.int32)
    if torch.device(device).type == "cuda":
        t = t.pin_memory().to(device, non_blocking=True)
    return t


def _cuda_f32(t: torch.Tensor) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError("veto_b200: tensors must live on a CUDA device (no CPU fallback)")
    return t.contiguous() if t.dtype == torch.float32 else t.float().contiguous()


# --------------------------------------------------------------------------------------------
# a1: pair enumeration
# --------------------------------------------------------------------------------------------
def pair_capacities(n_boxes: Sequence[int], max_pairs: int) -> List[int]:
    return [max(1, min(n * (n - 1), max_pairs)) for n in n_boxes]


def enumerate_pairs(n_boxes: Sequence[int], device, max_pairs: int = 2048, boxes: Optional[torch.Tensor] = None,
                    scores: Optional[torch.Tensor] = None, require_overlap: bool = False) -> List[torch.Tensor]:
    """RelationSampling.prepare_test_pairs (sampling.py:31-52): list of int64 [R_i,2] per image.

    `boxes` [N,4] xyxy / `scores` [N]: concatenated over images, needed for the IoU filter / the cap."""
    L.require_device()
    lib = L.load()
    B = len(n_boxes)
    if B == 0:
        return []
    caps = pair_capacities(n_boxes, max_pairs)
    total = sum(caps)
    pairs = torch.empty((total, 2), dtype=torch.int64, device=device)
    scratch = torch.empty(3 * (B + 1), dtype=torch.int32, device=device)
    counts = torch.empty(B, dtype=torch.int32, device=device) if require_overlap else None
    if boxes is not None:
        boxes = _cuda_f32(boxes)
    if scores is not None:
        scores = _cuda_f32(scores)
    with torch.cuda.device(pairs.device):
        L.check(lib.veto_pairs_enumerate(_i32(n_boxes), B, L.ptr(boxes), L.ptr(scores), int(bool(require_overlap)),
                                         int(max_pairs), pairs.data_ptr(), L.ptr(counts), scratch.data_ptr(),
                                         L.stream_ptr()), "veto_pairs_enumerate")
    if counts is None:
        n_valid = caps
    else:
        n_valid = counts.cpu().tolist()  # data-dependent sizes: the reference syncs here too (torch.nonzero)
    out, off = [], 0
    for cap, n in zip(caps, n_valid):
        out.append(pairs[off:off + n])
        off += cap
    return out


def globalize_pairs(rel_pair_idxs: Sequence[torch.Tensor], n_boxes: Sequence[int]) -> Tuple[torch.Tensor, torch.Tensor]:
    """roi_relation_predictors.py:4104-4115 on the device: global (subject, object) box indices, int32 [R]."""
    L.require_device()
    lib = L.load()
    device = rel_pair_idxs[0].device
    counts = [int(p.shape[0]) for p in rel_pair_idxs]
    R = sum(counts)
    pairs = rel_pair_idxs[0] if len(rel_pair_idxs) == 1 else torch.cat(list(rel_pair_idxs), 0)
    pairs = pairs.to(torch.int64).contiguous()
    subj = torch.empty(R, dtype=torch.int32, device=device)
    obj = torch.empty(R, dtype=torch.int32, device=device)
    if R == 0:
        return subj, obj
    rel_off = offsets_tensor(counts, device)
    box_off = offsets_tensor(n_boxes, device)
    with torch.cuda.device(device):
        L.check(lib.veto_pairs_globalize(pairs.data_ptr(), R, rel_off.data_ptr(), box_off.data_ptr(), len(counts),
                                         subj.data_ptr(), obj.data_ptr(), L.stream_ptr()), "veto_pairs_globalize")
    return subj, obj


# --------------------------------------------------------------------------------------------
# a3: ROIAlign
# --------------------------------------------------------------------------------------------
def roi_align_forward(inp: torch.Tensor, rois: torch.Tensor, spatial_scale: float, pooled_h: int, pooled_w: int,
                      sampling_ratio: int) -> torch.Tensor:
    """_C.roi_align_forward (pysgg/csrc/vision.cpp:11)."""
    L.require_device()
    inp, rois = _cuda_f32(inp), _cuda_f32(rois)
    B, C, H, W = inp.shape
    out = torch.empty((rois.shape[0], C, pooled_h, pooled_w), dtype=torch.float32, device=inp.device)
    with torch.cuda.device(inp.device):
        L.check(L.load().veto_roi_align_forward(inp.data_ptr(), B, C, H, W, rois.data_ptr(), rois.shape[0],
                                                float(spatial_scale), pooled_h, pooled_w, sampling_ratio,
                                                out.data_ptr(), L.stream_ptr()), "veto_roi_align_forward")
    return out


def roi_align_backward(grad: torch.Tensor, rois: torch.Tensor, spatial_scale: float, pooled_h: int, pooled_w: int,
                       batch: int, channels: int, height: int, width: int, sampling_ratio: int) -> torch.Tensor:
    """_C.roi_align_backward (pysgg/csrc/vision.cpp:12)."""
    L.require_device()
    grad, rois = _cuda_f32(grad), _cuda_f32(rois)
    gi = torch.empty((batch, channels, height, width), dtype=torch.float32, device=grad.device)
    with torch.cuda.device(grad.device):
        L.check(L.load().veto_roi_align_backward(grad.data_ptr(), rois.data_ptr(), rois.shape[0], float(spatial_scale),
                                                 pooled_h, pooled_w, batch, channels, height, width, sampling_ratio,
                                                 gi.data_ptr(), L.stream_ptr()), "veto_roi_align_backward")
    return gi


def roi_gather(feats: Sequence[torch.Tensor], depth: torch.Tensor, boxes: torch.Tensor, n_boxes: Sequence[int],
               scales: Sequence[float], depth_scale: float, pool: int = 8, sampling_ratio: int = 2,
               k_min: int = 2, k_max: int = 5, return_levels: bool = False):
    """Pooler.forward depth + RGB branch (poolers.py:109-171) in one launch: (x_2d, d_2d) [N,C,pool,pool]."""
    L.require_device()
    feats = [_cuda_f32(f) for f in feats]
    depth, boxes = _cuda_f32(depth), _cuda_f32(boxes)
    n_levels = len(feats)
    Bt, C = feats[0].shape[:2]
    N = boxes.shape[0]
    dev = boxes.device
    x2d = torch.empty((N, C, pool, pool), dtype=torch.float32, device=dev)
    d2d = torch.empty((N, C, pool, pool), dtype=torch.float32, device=dev)
    levels = torch.empty(N, dtype=torch.int32, device=dev)   # always: the library then computes the FPN level once per box
    if N:
        box_off = offsets_tensor(n_boxes, dev)
        fp = (ctypes.c_void_p * n_levels)(*[f.data_ptr() for f in feats])
        fh = _i32([f.shape[2] for f in feats])
        fw = _i32([f.shape[3] for f in feats])
        sc = (ctypes.c_float * n_levels)(*[float(s) for s in scales])
        with torch.cuda.device(dev):
            L.check(L.load().veto_roi_gather_forward(fp, fh, fw, sc, n_levels, k_min, k_max, depth.data_ptr(),
                                                     depth.shape[2], depth.shape[3], float(depth_scale), Bt, C,
                                                     boxes.data_ptr(), box_off.data_ptr(), len(n_boxes), N, pool,
                                                     sampling_ratio, x2d.data_ptr(), d2d.data_ptr(), L.ptr(levels),
                                                     L.stream_ptr()), "veto_roi_gather_forward")
    return (x2d, d2d, levels) if return_levels else (x2d, d2d)


# --------------------------------------------------------------------------------------------
# a5-a10: relation head
# --------------------------------------------------------------------------------------------
def make_config(num_obj: int, num_out: int, precision: str = "bf16x3", layers: int = 6, dim: int = 576, heads: int = 6,
                mlp_dim: int = 1152, channels: int = 256, pool: int = 8, patch: int = 2) -> L.VetoConfig:
    if precision not in L.PRECISIONS:
        raise ValueError(f"precision must be one of {sorted(L.PRECISIONS)}, got {precision!r}")
    return L.VetoConfig(dim, layers, heads, mlp_dim, channels, pool, patch, num_obj, num_out, L.PRECISIONS[precision])


# state_dict key (relative to the predictor / Ensemble module) -> veto_weights field
_SCALAR_KEYS = {
    "obj_embed": "obj_embed.weight",
    "class_proj_w": "class_projection.0.weight", "class_proj_b": "class_projection.0.bias",
    "bn_weight": "pos_embed.0.weight", "bn_bias": "pos_embed.0.bias",
    "bn_mean": "pos_embed.0.running_mean", "bn_var": "pos_embed.0.running_var",
    "pos_w": "pos_embed.1.weight", "pos_b": "pos_embed.1.bias",
    "loc_proj_w": "location_projection.0.weight", "loc_proj_b": "location_projection.0.bias",
    "cls_token": "fusion_transformer.transformer.cls_token",
    "pos_embedding": "fusion_transformer.transformer.pos_embedding",
    "proj_d_w": "fusion_transformer.transformer.patch_embed.proj_d.weight",
    "proj_d_b": "fusion_transformer.transformer.patch_embed.proj_d.bias",
    "proj_v_w": "fusion_transformer.transformer.patch_embed.proj_v.weight",
    "proj_v_b": "fusion_transformer.transformer.patch_embed.proj_v.bias",
}
_LAYER_KEYS = {
    "ln1_w": "0.norm.weight", "ln1_b": "0.norm.bias", "qkv_w": "0.fn.to_qkv.weight",
    "out_w": "0.fn.to_out.0.weight", "out_b": "0.fn.to_out.0.bias",
    "ln2_w": "1.norm.weight", "ln2_b": "1.norm.bias",
    "ff1_w": "1.fn.net.0.weight", "ff1_b": "1.fn.net.0.bias", "ff2_w": "1.fn.net.3.weight", "ff2_b": "1.fn.net.3.bias",
}


class PackedWeights:
    """veto_weights + the packed device buffer for one (state, precision).  Keeps the source tensors alive."""

    def __init__(self, cfg: L.VetoConfig, tensors: Dict[str, torch.Tensor], rel_out_w: torch.Tensor,
                 rel_out_b: torch.Tensor):
        L.require_device()
        lib = L.load()
        self.cfg = cfg
        keep = []

        def dev(t):
            t = _cuda_f32(t.detach())
            keep.append(t)
            return t.data_ptr()

        w = L.VetoWeights()
        for field, key in _SCALAR_KEYS.items():
            setattr(w, field, dev(tensors[key]))
        for field, key in _LAYER_KEYS.items():
            arr = getattr(w, field)
            for i in range(cfg.layers):
                arr[i] = dev(tensors[f"fusion_transformer.transformer.layers.{i}.{key}"])
        w.rel_out_w = dev(rel_out_w)
        w.rel_out_b = dev(rel_out_b)
        self.struct = w
        self._keep = keep
        self.device = keep[0].device
        nbytes = lib.veto_packed_bytes(ctypes.byref(cfg))
        if nbytes == 0:
            L.check(-3, "veto_packed_bytes")
        self.packed = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            L.check(lib.veto_pack_weights(ctypes.byref(cfg), ctypes.byref(w), self.packed.data_ptr(), nbytes,
                                          L.stream_ptr()), "veto_pack_weights")


_workspaces: Dict[Tuple, torch.Tensor] = {}


def _workspace(device, nbytes: int) -> torch.Tensor:
    key = (str(device),)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = None
        _workspaces.pop(key, None)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


last_launch_count = 0


def relation_forward(pw: PackedWeights, boxes: torch.Tensor, roi_rgb: torch.Tensor, roi_depth: torch.Tensor,
                     subj: torch.Tensor, obj: torch.Tensor, labels: Optional[torch.Tensor] = None,
                     obj_logits: Optional[torch.Tensor] = None, freq_bias: Optional[torch.Tensor] = None,
                     chunk_pairs: int = 0, return_features: bool = False, return_tokens: bool = False):
    """Relation logits [R, num_out] of VETOPredictor / Ensemble (eval) for global pair indices subj/obj."""
    global last_launch_count
    lib = L.load()
    cfg = pw.cfg
    dev = pw.device
    boxes, roi_rgb, roi_depth = _cuda_f32(boxes), _cuda_f32(roi_rgb), _cuda_f32(roi_depth)
    N, R = boxes.shape[0], subj.shape[0]
    if tuple(roi_rgb.shape) != (N, 256, 8, 8) or tuple(roi_depth.shape) != (N, 256, 8, 8):
        raise RuntimeError(f"roi features must be [{N},256,8,8], got {tuple(roi_rgb.shape)} / {tuple(roi_depth.shape)}")
    subj = subj.to(torch.int32).contiguous()
    obj = obj.to(torch.int32).contiguous()
    if labels is not None:
        labels = labels.to(torch.int64).contiguous()
    if obj_logits is not None:
        obj_logits = _cuda_f32(obj_logits)
    logits = torch.empty((R, cfg.num_out), dtype=torch.float32, device=dev)
    feats = torch.empty((R, T_DIM), dtype=torch.float32, device=dev) if return_features else None
    toks = torch.empty((R, N_TOKENS, T_DIM), dtype=torch.float32, device=dev) if return_tokens else None
    if R and N:
        nbytes = lib.veto_workspace_bytes(ctypes.byref(cfg), N, R, chunk_pairs)
        ws = _workspace(dev, nbytes)
        vin = L.VetoInputs(N, R, boxes.data_ptr(), L.ptr(labels), L.ptr(obj_logits), roi_rgb.data_ptr(),
                           roi_depth.data_ptr(), subj.data_ptr(), obj.data_ptr(), L.ptr(freq_bias))
        vout = L.VetoOutputs(logits.data_ptr(), L.ptr(feats), L.ptr(toks))
        with torch.cuda.device(dev):
            L.check(lib.veto_relation_forward(ctypes.byref(cfg), ctypes.byref(pw.struct), pw.packed.data_ptr(),
                                              ctypes.byref(vin), ctypes.byref(vout), ws.data_ptr(), ws.numel(),
                                              chunk_pairs, L.stream_ptr()), "veto_relation_forward")
        last_launch_count = int(lib.veto_last_launch_count())
    out = [logits]
    if return_features:
        out.append(feats)
    if return_tokens:
        out.append(toks)
    return out[0] if len(out) == 1 else tuple(out)


# --------------------------------------------------------------------------------------------
# a11: training step (forward in train() mode + loss + backward)
# --------------------------------------------------------------------------------------------
def grad_fields(n_layers: int) -> List[Tuple[str, Optional[int], str]]:
    """(veto_grads field, layer or None, state_dict key relative to the trunk) of every parameter the training
    step produces a gradient for, in a fixed order (the layout of the flat gradient buffer)."""
    out = []
    for field, key in _SCALAR_KEYS.items():
        if field in ("bn_mean", "bn_var"):
            continue  # running statistics: buffers, no gradient
        out.append((field, None, key))
    for i in range(n_layers):
        for field, key in _LAYER_KEYS.items():
            out.append((field, i, f"fusion_transformer.transformer.layers.{i}.{key}"))
    return out


def relation_train_step(pw: PackedWeights, boxes: torch.Tensor, roi_rgb: torch.Tensor, roi_depth: torch.Tensor,
                        subj: torch.Tensor, obj: torch.Tensor, rel_labels: torch.Tensor, rel_counts: Sequence[int],
                        n_boxes: Sequence[int], grads: Dict[str, torch.Tensor], rel_out_w_grad: torch.Tensor,
                        rel_out_b_grad: torch.Tensor, labels: Optional[torch.Tensor] = None,
                        obj_logits: Optional[torch.Tensor] = None, class_weight: Optional[torch.Tensor] = None,
                        p_pos: float = 0.1, p_emb: float = 0.35, p_attn: float = 0.35, seed: int = 0,
                        bn_momentum: float = 0.001, bn_running_mean: Optional[torch.Tensor] = None,
                        bn_running_var: Optional[torch.Tensor] = None, want_roi_depth_grad: bool = True,
                        want_roi_rgb_grad: bool = False, return_logits: bool = False,
                        head_sizes: Optional[Sequence[int]] = None, head_labels: Optional[torch.Tensor] = None):
    """veto_relation_train_step: rel_loss of VETOPredictor.forward in train() mode and its gradients.

    `grads` maps the trunk's state_dict keys (grad_fields) to fp32 CUDA tensors of the parameter shapes; they are
    OVERWRITTEN.  Returns (loss [1], grad_roi_depth or None, grad_roi_rgb or None, logits or None).

    MEET group heads: `head_sizes` = the n_k + 2 columns of each head (their sum = cfg.num_out) and `head_labels`
    int64 [n_heads, R] (group-local label, -1 = pair not sampled into that head's loss); loss is then [n_heads] and
    the gradients are those of the sum of the head losses; `rel_labels` / `class_weight` are ignored."""
    global last_launch_count
    lib = L.load()
    cfg = pw.cfg
    dev = pw.device
    boxes, roi_rgb, roi_depth = _cuda_f32(boxes), _cuda_f32(roi_rgb), _cuda_f32(roi_depth)
    N, R = boxes.shape[0], subj.shape[0]
    if N == 0 or R == 0:
        raise RuntimeError("a training step needs at least one box and one pair")
    if tuple(roi_rgb.shape) != (N, 256, 8, 8) or tuple(roi_depth.shape) != (N, 256, 8, 8):
        raise RuntimeError(f"roi features must be [{N},256,8,8], got {tuple(roi_rgb.shape)} / {tuple(roi_depth.shape)}")
    n_heads = len(head_sizes) if head_sizes is not None else 0
    if n_heads > 1:
        if sum(head_sizes) != cfg.num_out or head_labels is None or tuple(head_labels.shape) != (n_heads, R):
            raise RuntimeError("head_sizes must sum to num_out and head_labels must be [n_heads, R]")
        head_labels = head_labels.to(device=dev, dtype=torch.int64).contiguous()
        head_off = (ctypes.c_int32 * (n_heads + 1))(*([0] + list(itertools.accumulate(int(n) for n in head_sizes))))
        rel_labels = head_labels[0]
        class_weight = None
    if sum(rel_counts) != R or sum(n_boxes) != N or rel_labels.shape[0] != R:
        raise RuntimeError("rel_counts / n_boxes / rel_labels do not match the pair and box tensors")
    subj = subj.to(torch.int32).contiguous()
    obj = obj.to(torch.int32).contiguous()
    rel_labels = rel_labels.to(torch.int64).contiguous()
    if labels is not None:
        labels = labels.to(torch.int64).contiguous()
    if obj_logits is not None:
        obj_logits = _cuda_f32(obj_logits)
    if class_weight is not None:
        class_weight = _cuda_f32(class_weight)
    rel_off = offsets_tensor(rel_counts, dev)
    box_off = offsets_tensor(n_boxes, dev)
    g = L.VetoGrads()
    for field, layer, key in grad_fields(cfg.layers):
        t = grads[key]
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            raise RuntimeError(f"gradient buffer of {key} must be a contiguous fp32 CUDA tensor")
        if layer is None:
            setattr(g, field, t.data_ptr())
        else:
            getattr(g, field)[layer] = t.data_ptr()
    g.rel_out_w, g.rel_out_b = rel_out_w_grad.data_ptr(), rel_out_b_grad.data_ptr()
    loss = torch.empty(max(1, n_heads), dtype=torch.float32, device=dev)
    logits = torch.empty((R, cfg.num_out), dtype=torch.float32, device=dev) if return_logits else None
    g_depth = torch.empty_like(roi_depth) if want_roi_depth_grad else None
    g_rgb = torch.empty_like(roi_rgb) if want_roi_rgb_grad else None
    nbytes = lib.veto_train_workspace_bytes(ctypes.byref(cfg), N, R)
    ws = _workspace(dev, nbytes)
    vin = L.VetoInputs(N, R, boxes.data_ptr(), L.ptr(labels), L.ptr(obj_logits), roi_rgb.data_ptr(),
                       roi_depth.data_ptr(), subj.data_ptr(), obj.data_ptr(), None)
    tin = L.VetoTrainInputs(rel_labels.data_ptr(), L.ptr(class_weight), rel_off.data_ptr(), box_off.data_ptr(),
                            len(rel_counts), float(p_pos), float(p_emb), float(p_attn), int(seed) & (2 ** 64 - 1),
                            float(bn_momentum), L.ptr(bn_running_mean), L.ptr(bn_running_var),
                            n_heads if n_heads > 1 else 0,
                            ctypes.cast(head_off, ctypes.c_void_p) if n_heads > 1 else None,
                            head_labels.data_ptr() if n_heads > 1 else None)
    tout = L.VetoTrainOutputs(loss.data_ptr(), L.ptr(logits), L.ptr(g_depth), L.ptr(g_rgb))
    with torch.cuda.device(dev):
        L.check(lib.veto_relation_train_step(ctypes.byref(cfg), ctypes.byref(pw.struct), pw.packed.data_ptr(),
                                             ctypes.byref(vin), ctypes.byref(tin), ctypes.byref(g), ctypes.byref(tout),
                                             ws.data_ptr(), ws.numel(), L.stream_ptr()), "veto_relation_train_step")
    last_launch_count = int(lib.veto_last_launch_count())
    return loss, g_depth, g_rgb, logits


def dropout_keep_mask(seed: int, sub: int, n: int, p: float):
    """The keep mask (numpy bool [n]) the library uses for stream `sub` of a step seeded `seed` (csrc/common.cuh
    drop_hash / train_api.cu sub_seed): sub 1 = pos_embed dropout over [N,128], 2 = token dropout over [R,19,576],
    16 + l = to_out dropout of layer l over [R*19,576].  For tests that reproduce a dropped step on the oracle."""
    import numpy as np
    M64 = np.uint64(0xFFFFFFFFFFFFFFFF)

    def mix(seed64, group):
        with np.errstate(over="ignore"):
            z = (seed64 + (group + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)) & M64
            z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & M64
            z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & M64
            return z ^ (z >> np.uint64(31))

    s = mix(np.uint64(seed & (2 ** 64 - 1)) ^ np.uint64(0xA5A5A5A5A5A5A5A5), np.uint64(sub))
    e = np.arange(n, dtype=np.uint64)
    h = mix(s, e >> np.uint64(2))
    field = (h >> (np.uint64(16) * (e & np.uint64(3)))) & np.uint64(0xFFFF)
    thr = np.uint64(int(p * 65536.0 + 0.5))
    return field >= thr if p > 0 else np.ones(n, dtype=bool)


# --------------------------------------------------------------------------------------------
# f2: training-time relation sampling (ground-truth boxes)
# --------------------------------------------------------------------------------------------
def relsample_gtbox(rel_matrices: Sequence[torch.Tensor], batch_size_per_image: int, num_pos_per_image: int, seed: int):
    """RelationSampling.gtbox_relsample (sampling.py:54-107) for a batch in one launch.  rel_matrices: per image the
    target's [n, n] "relation" matrix.  Returns (pairs [B*batch,2] int64, labels [B*batch] int64, counts [B,2] int32
    (foreground rows, total rows) — image b's rows start at b*batch — and the per-image binary matrices)."""
    L.require_device()
    n_boxes = [int(m.shape[0]) for m in rel_matrices]
    dev = rel_matrices[0].device
    if any(m.dim() != 2 or m.shape[0] != m.shape[1] for m in rel_matrices):
        raise RuntimeError("relation matrices must be square")
    flat = torch.cat([m.reshape(-1) for m in rel_matrices]).to(torch.int64).contiguous()
    mat_off = offsets_tensor([n * n for n in n_boxes], dev)
    box_off = offsets_tensor(n_boxes, dev)
    B = len(n_boxes)
    pairs = torch.zeros((B * batch_size_per_image, 2), dtype=torch.int64, device=dev)
    labels = torch.zeros(B * batch_size_per_image, dtype=torch.int64, device=dev)
    counts = torch.zeros((B, 2), dtype=torch.int32, device=dev)
    binary = torch.empty_like(flat)
    nb = (ctypes.c_int32 * B)(*n_boxes)
    with torch.cuda.device(dev):
        L.check(L.load().veto_relsample_gtbox(flat.data_ptr(), mat_off.data_ptr(), box_off.data_ptr(), nb, B,
                                              int(batch_size_per_image), int(num_pos_per_image), int(seed) & (2 ** 64 - 1),
                                              pairs.data_ptr(), labels.data_ptr(), counts.data_ptr(), binary.data_ptr(),
                                              L.stream_ptr()), "veto_relsample_gtbox")
    binaries, off = [], 0
    for n in n_boxes:
        binaries.append(binary[off:off + n * n].view(n, n))
        off += n * n
    return pairs, labels, counts, binaries


def relsample_detect(prp_boxes: Sequence[torch.Tensor], prp_labels: Sequence[torch.Tensor], prp_scores: Sequence[torch.Tensor],
                     tgt_boxes: Sequence[torch.Tensor], tgt_labels: Sequence[torch.Tensor], tgt_rels: Sequence[torch.Tensor],
                     fg_thres: float, require_overlap: bool, num_sample_per_gt_rel: int, batch_size_per_image: int,
                     num_pos_per_image: int, seed: int):
    """RelationSampling.detect_relsample (sampling.py:109-309) for a batch in one launch; per-image lists in.
    Returns (triplets [B*batch,3], corrsp [B*batch], counts [B,2] int32 (fg rows, total rows), binaries (per image
    [P,P]), locating (per image [P])); image b's rows start at b*batch."""
    L.require_device()
    dev = prp_boxes[0].device
    n_prp = [int(b.shape[0]) for b in prp_boxes]
    n_tgt = [int(b.shape[0]) for b in tgt_boxes]
    B = len(n_prp)
    pb = _cuda_f32(torch.cat(list(prp_boxes), 0)).reshape(-1, 4)
    tb = _cuda_f32(torch.cat(list(tgt_boxes), 0)).reshape(-1, 4)
    pl = torch.cat(list(prp_labels), 0).to(torch.int64).contiguous()
    tl = torch.cat(list(tgt_labels), 0).to(torch.int64).contiguous()
    ps = _cuda_f32(torch.cat(list(prp_scores), 0))
    rel = torch.cat([r.reshape(-1) for r in tgt_rels]).to(torch.int64).contiguous()
    p_off, t_off = offsets_tensor(n_prp, dev), offsets_tensor(n_tgt, dev)
    r_off = offsets_tensor([n * n for n in n_tgt], dev)
    b_off = offsets_tensor([n * n for n in n_prp], dev)
    trip = torch.zeros((B * batch_size_per_image, 3), dtype=torch.int64, device=dev)
    corr = torch.full((B * batch_size_per_image,), -1, dtype=torch.int64, device=dev)
    counts = torch.zeros((B, 2), dtype=torch.int32, device=dev)
    binary = torch.zeros(max(1, sum(n * n for n in n_prp)), dtype=torch.int64, device=dev)
    locating = torch.zeros(max(1, sum(n_prp)), dtype=torch.float32, device=dev)
    np_h = (ctypes.c_int32 * B)(*n_prp)
    nt_h = (ctypes.c_int32 * B)(*n_tgt)
    with torch.cuda.device(dev):
        L.check(L.load().veto_relsample_detect(
            pb.data_ptr(), pl.data_ptr(), ps.data_ptr(), p_off.data_ptr(), tb.data_ptr(), tl.data_ptr(), t_off.data_ptr(),
            rel.data_ptr(), r_off.data_ptr(), b_off.data_ptr(), np_h, nt_h, B, float(fg_thres), int(bool(require_overlap)),
            int(num_sample_per_gt_rel), int(batch_size_per_image), int(num_pos_per_image), int(seed) & (2 ** 64 - 1),
            trip.data_ptr(), corr.data_ptr(), counts.data_ptr(), binary.data_ptr(), locating.data_ptr(), L.stream_ptr()),
            "veto_relsample_detect")
    binaries, locs, bo, po = [], [], 0, 0
    for n in n_prp:
        binaries.append(binary[bo:bo + n * n].view(n, n))
        locs.append(locating[po:po + n])
        bo += n * n
        po += n
    return trip, corr, counts, binaries, locs


# --------------------------------------------------------------------------------------------
# a10: MEET's per-class NMS label assignment (SGDet test)
# --------------------------------------------------------------------------------------------
ZERO_MODES = {"rand_insert": 0, "rand_choose": 1, "all_include": 2}


def meet_group_labels(rel_labels: torch.Tensor, incre_idx: torch.Tensor, rates: torch.Tensor, local_label: torch.Tensor,
                      zero_mode: str, seed: int, draws: Optional[torch.Tensor] = None,
                      bg_heads: Optional[torch.Tensor] = None) -> torch.Tensor:
    """cur_chosen_matrix + per-head relabelling of VETOPredictor_MEET's training branch on the device
    (roi_relation_predictors.py:3940-3969, 3812-3821): int64 [n_groups, R] head-local labels, -1 = pair not in the head's
    loss.  Tables (device): incre_idx int32 [num_rel], rates float64 [G, num_rel], local_label int32 [G, num_rel]."""
    L.require_device()
    if zero_mode not in ZERO_MODES:
        raise ValueError(f"ZERO_LABEL_PADDING_MODE must be one of {sorted(ZERO_MODES)}, got {zero_mode!r}")
    rel_labels = rel_labels.to(torch.int64).contiguous()
    G, num_rel = rates.shape
    if (rates.dtype != torch.float64 or incre_idx.dtype != torch.int32 or local_label.dtype != torch.int32
            or tuple(local_label.shape) != (G, num_rel) or incre_idx.numel() != num_rel):
        raise RuntimeError("meet_group_labels: tables must be float64 [G,num_rel] / int32 [num_rel] / int32 [G,num_rel]")
    R = rel_labels.numel()
    out = torch.empty((G, R), dtype=torch.int64, device=rel_labels.device)
    if R:
        if draws is not None:
            draws = draws.to(device=rel_labels.device, dtype=torch.float64).contiguous()
        if bg_heads is not None:
            bg_heads = bg_heads.to(device=rel_labels.device, dtype=torch.int32).contiguous()
        with torch.cuda.device(rel_labels.device):
            L.check(L.load().veto_meet_group_labels(rel_labels.data_ptr(), R, incre_idx.data_ptr(), rates.data_ptr(),
                                                    local_label.data_ptr(), G, num_rel, ZERO_MODES[zero_mode],
                                                    int(seed) & (2 ** 64 - 1), L.ptr(draws), L.ptr(bg_heads), out.data_ptr(),
                                                    L.stream_ptr()), "veto_meet_group_labels")
    return out


def obj_nms_per_cls(scores: torch.Tensor, boxes_per_cls: torch.Tensor, n_boxes: Sequence[int], thresh: float,
                    late_nms: bool = False) -> torch.Tensor:
    """Ensemble.nms_per_cls (roi_relation_predictors.py:3855-3874), or with late_nms obj_prediction_nms
    (relation_head/utils_relation.py:94-128): scores [N,num_obj] fp32 (softmax of the object distribution),
    boxes_per_cls [N,num_obj,4]; returns int64 labels [N].  One kernel launch for the whole batch."""
    L.require_device()
    scores, boxes_per_cls = _cuda_f32(scores), _cuda_f32(boxes_per_cls)
    N, num_obj = scores.shape
    if tuple(boxes_per_cls.shape) != (N, num_obj, 4) or sum(n_boxes) != N:
        raise RuntimeError(f"boxes_per_cls must be [{N},{num_obj},4] and n_boxes must sum to {N}")
    dev = scores.device
    out = torch.zeros(N, dtype=torch.int64, device=dev)
    if N:
        box_off = offsets_tensor(n_boxes, dev)
        nb = (ctypes.c_int32 * len(n_boxes))(*[int(n) for n in n_boxes])
        with torch.cuda.device(dev):
            L.check(L.load().veto_obj_nms_per_cls(scores.data_ptr(), boxes_per_cls.data_ptr(), box_off.data_ptr(), nb,
                                                  len(n_boxes), num_obj, float(thresh), int(bool(late_nms)), out.data_ptr(),
                                                  L.stream_ptr()),
                    "veto_obj_nms_per_cls")
    return out


# --------------------------------------------------------------------------------------------
# a12: post-processing
# --------------------------------------------------------------------------------------------
def postprocess(rel_logits: torch.Tensor, pairs: torch.Tensor, obj_scores: torch.Tensor, rel_counts: Sequence[int],
                n_boxes: Sequence[int]):
    """PostProcessor vanilla branch (inference.py:398-453) for a whole batch: returns image-segmented
    (sorted pairs [R,2], class probabilities [R,C], labels [R], triple scores [R])."""
    L.require_device()
    rel_logits, obj_scores = _cuda_f32(rel_logits), _cuda_f32(obj_scores)
    pairs = pairs.to(torch.int64).contiguous()
    R, C = rel_logits.shape
    dev = rel_logits.device
    pairs_o = torch.empty_like(pairs)
    probs_o = torch.empty_like(rel_logits)
    labels_o = torch.empty(R, dtype=torch.int64, device=dev)
    triple_o = torch.empty(R, dtype=torch.float32, device=dev)
    if R:
        rel_off = offsets_tensor(rel_counts, dev)
        box_off = offsets_tensor(n_boxes, dev)
        with torch.cuda.device(dev):
            L.check(L.load().veto_postprocess(rel_logits.data_ptr(), C, pairs.data_ptr(), obj_scores.data_ptr(),
                                              rel_off.data_ptr(), box_off.data_ptr(), len(rel_counts), R,
                                              pairs_o.data_ptr(), probs_o.data_ptr(), labels_o.data_ptr(),
                                              triple_o.data_ptr(), L.stream_ptr()), "veto_postprocess")
    return pairs_o, probs_o, labels_o, triple_o


def postprocess_meet(group_logits: torch.Tensor, head_sizes: Sequence[int], col_map: Sequence[int], num_rel: int,
                     pairs: torch.Tensor, obj_scores: torch.Tensor, rel_counts: Sequence[int], n_boxes: Sequence[int]):
    """PostProcessor MEET 'ensemble' branch (inference.py:284-397) for a batch: group_logits [R, sum(n_k+2)], col_map =
    global predicate id of every concatenated column.  Returns image-segmented (G*R_i rows per image): sorted pairs
    int64 [G*R,2], probabilities [G*R,num_rel] in global columns, head-local labels [G*R], triple scores [G*R]."""
    L.require_device()
    group_logits, obj_scores = _cuda_f32(group_logits), _cuda_f32(obj_scores)
    pairs = pairs.to(torch.int64).contiguous()
    R, Ct = group_logits.shape
    G = len(head_sizes)
    dev = group_logits.device
    if sum(head_sizes) != Ct or len(col_map) != Ct or sum(rel_counts) != R:
        raise RuntimeError("head_sizes / col_map / rel_counts do not match the group logits")
    pairs_o = torch.empty((G * R, 2), dtype=torch.int64, device=dev)
    probs_o = torch.empty((G * R, num_rel), dtype=torch.float32, device=dev)
    labels_o = torch.empty(G * R, dtype=torch.int64, device=dev)
    triple_o = torch.empty(G * R, dtype=torch.float32, device=dev)
    if R:
        rel_off = offsets_tensor(rel_counts, dev)
        box_off = offsets_tensor(n_boxes, dev)
        head_off = torch.tensor([0] + list(itertools.accumulate(int(n) for n in head_sizes)), dtype=torch.int32, device=dev)
        cmap = torch.tensor([int(c) for c in col_map], dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            L.check(L.load().veto_postprocess_meet(group_logits.data_ptr(), Ct, head_off.data_ptr(), G, cmap.data_ptr(),
                                                   num_rel, pairs.data_ptr(), obj_scores.data_ptr(), rel_off.data_ptr(),
                                                   box_off.data_ptr(), len(rel_counts), R, pairs_o.data_ptr(),
                                                   probs_o.data_ptr(), labels_o.data_ptr(), triple_o.data_ptr(),
                                                   L.stream_ptr()), "veto_postprocess_meet")
    return pairs_o, probs_o, labels_o, triple_o


def postprocess_meet_vote(group_logits: torch.Tensor, head_sizes: Sequence[int], col_map: Sequence[int], num_rel: int,
                          consensus: bool, pairs: torch.Tensor, obj_scores: torch.Tensor, rel_counts: Sequence[int],
                          n_boxes: Sequence[int]):
    """PostProcessor MEET EXPERT_GROUP branch (inference.py:93-283): group_logits [R, 3 * sum(n_k+2)] expert-major,
    head_sizes / col_map for all 3*G heads.  Returns (pairs [G*R,2], probs [G*R,num_rel], labels [G*R], triple [G*R],
    counts [B] int32 — survivors per image; image b's rows start at G * (pairs before it))."""
    L.require_device()
    group_logits, obj_scores = _cuda_f32(group_logits), _cuda_f32(obj_scores)
    pairs = pairs.to(torch.int64).contiguous()
    R, Ct = group_logits.shape
    if len(head_sizes) % 3 or sum(head_sizes) != Ct or len(col_map) != Ct or sum(rel_counts) != R:
        raise RuntimeError("head_sizes (3 experts x G groups) / col_map / rel_counts do not match the group logits")
    G = len(head_sizes) // 3
    if any(head_sizes[e * G + j] != head_sizes[j] for e in range(3) for j in range(G)):
        raise RuntimeError("the three experts of a group must have the same number of outputs")
    dev = group_logits.device
    pairs_o = torch.zeros((G * R, 2), dtype=torch.int64, device=dev)
    probs_o = torch.zeros((G * R, num_rel), dtype=torch.float32, device=dev)
    labels_o = torch.zeros(G * R, dtype=torch.int64, device=dev)
    triple_o = torch.zeros(G * R, dtype=torch.float32, device=dev)
    counts = torch.zeros(len(rel_counts), dtype=torch.int32, device=dev)
    if R:
        rel_off = offsets_tensor(rel_counts, dev)
        box_off = offsets_tensor(n_boxes, dev)
        head_off = torch.tensor([0] + list(itertools.accumulate(int(n) for n in head_sizes)), dtype=torch.int32, device=dev)
        cmap = torch.tensor([int(c) for c in col_map], dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            L.check(L.load().veto_postprocess_meet_vote(
                group_logits.data_ptr(), Ct, head_off.data_ptr(), G, cmap.data_ptr(), num_rel, int(bool(consensus)),
                pairs.data_ptr(), obj_scores.data_ptr(), rel_off.data_ptr(), box_off.data_ptr(), len(rel_counts), R,
                pairs_o.data_ptr(), probs_o.data_ptr(), labels_o.data_ptr(), triple_o.data_ptr(), counts.data_ptr(),
                L.stream_ptr()), "veto_postprocess_meet_vote")
    return pairs_o, probs_o, labels_o, triple_o, counts


# --------------------------------------------------------------------------------------------
# f4: evaluation triplet matching
# --------------------------------------------------------------------------------------------
def sgg_match(gt_triplets: torch.Tensor, gt_boxes: torch.Tensor, gt_counts: Sequence[int], pred_triplets: torch.Tensor,
              pred_boxes: torch.Tensor, pred_counts: Sequence[int], iou_thres: float = 0.5):
    """_compute_pred_matches (sgg_eval.py:77-117) for a batch: triplets int64 [.,3], boxes fp32 [.,8] (subject xyxy,
    object xyxy), per-image row counts.  Returns (first_match int32 [G] — rank of the first matching prediction within
    the image, INT32_MAX = none —, pred_hits int32 [P] — ground-truth triplets matched by each prediction)."""
    L.require_device()
    gt_boxes, pred_boxes = _cuda_f32(gt_boxes), _cuda_f32(pred_boxes)
    gt_triplets = gt_triplets.to(torch.int64).contiguous()
    pred_triplets = pred_triplets.to(torch.int64).contiguous()
    dev = gt_boxes.device
    G, P = gt_triplets.shape[0], pred_triplets.shape[0]
    if sum(gt_counts) != G or sum(pred_counts) != P or len(gt_counts) != len(pred_counts):
        raise RuntimeError("per-image counts do not match the triplet rows")
    if tuple(gt_boxes.shape) != (G, 8) or tuple(pred_boxes.shape) != (P, 8):
        raise RuntimeError("triplet boxes must be [rows, 8]")
    first = torch.full((G,), 2 ** 31 - 1, dtype=torch.int32, device=dev)
    hits = torch.zeros(P, dtype=torch.int32, device=dev)
    if G and P:
        g_off, p_off = offsets_tensor(gt_counts, dev), offsets_tensor(pred_counts, dev)
        with torch.cuda.device(dev):
            L.check(L.load().veto_sgg_match(gt_triplets.data_ptr(), gt_boxes.data_ptr(), g_off.data_ptr(),
                                            pred_triplets.data_ptr(), pred_boxes.data_ptr(), p_off.data_ptr(),
                                            len(gt_counts), float(iou_thres), first.data_ptr(), hits.data_ptr(),
                                            L.stream_ptr()), "veto_sgg_match")
    return first, hits


# --------------------------------------------------------------------------------------------
# f3: the depth backbone (ResNetDepth, backbone/resnet_depth.py:11-47)
# --------------------------------------------------------------------------------------------
def depth_backbone_out_size(height: int, width: int):
    h, w = ctypes.c_int(0), ctypes.c_int(0)
    L.load().veto_depth_backbone_out_size(int(height), int(width), ctypes.byref(h), ctypes.byref(w))
    return h.value, w.value


def _depth_weight_struct(convs, bn_w, bn_b, bn_mean, bn_var):
    W = L.VetoDepthWeights()
    for i in range(L.DEPTH_CONVS):
        W.conv_w[i], W.bn_w[i], W.bn_b[i] = convs[i].data_ptr(), bn_w[i].data_ptr(), bn_b[i].data_ptr()
        W.bn_mean[i], W.bn_var[i] = bn_mean[i].data_ptr(), bn_var[i].data_ptr()
    return W


def depth_backbone_forward(depth: torch.Tensor, convs, bn_w, bn_b, bn_mean, bn_var, training: bool, momentum: float = 0.1,
                           precision: str = "bf16x3"):
    """ResNetDepth.forward: depth [B,1,H,W] -> [B,256,H/16,W/16] (NCHW).  The parameter lists are in the module order
    of include/veto_b200.h (15 convolutions, each with its BatchNorm2d); in training mode the running statistics are
    updated in place and the returned workspace holds what ``depth_backbone_backward`` needs."""
    L.require_device()
    depth = _cuda_f32(depth)
    if depth.dim() != 4 or depth.shape[1] != 1:
        raise RuntimeError("depth images must be [B,1,H,W]")
    if len(convs) != L.DEPTH_CONVS:
        raise RuntimeError("the depth backbone has %d convolutions" % L.DEPTH_CONVS)
    B, _, H, W_ = depth.shape
    prec = L.PRECISIONS[L.TRAIN_PRECISION[precision]]
    lib = L.load()
    nbytes = lib.veto_depth_backbone_workspace_bytes(prec, B, H, W_, int(training))
    if nbytes == 0:
        L.check(-1, "veto_depth_backbone_workspace_bytes")
    ws = torch.empty(nbytes, dtype=torch.uint8, device=depth.device)
    oh, ow = depth_backbone_out_size(H, W_)
    out = torch.empty(B, 256, oh, ow, dtype=torch.float32, device=depth.device)
    tensors = [[_cuda_f32(t) for t in group] for group in (convs, bn_w, bn_b)]
    for group in (bn_mean, bn_var):      # updated in place: must already be contiguous fp32 on the device
        for t in group:
            if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
                raise RuntimeError("BatchNorm running statistics must be contiguous fp32 CUDA tensors")
    Wst = _depth_weight_struct(tensors[0], tensors[1], tensors[2], bn_mean, bn_var)
    with torch.cuda.device(depth.device):
        L.check(lib.veto_depth_backbone_forward(prec, ctypes.byref(Wst), depth.data_ptr(), B, H, W_, int(training),
                                                float(momentum), out.data_ptr(), ws.data_ptr(), nbytes, L.stream_ptr()),
                "veto_depth_backbone_forward")
    return out, ws


def depth_backbone_backward(grad_out: torch.Tensor, depth_shape, convs, bn_w, bn_b, bn_mean, bn_var, ws: torch.Tensor,
                            precision: str = "bf16x3"):
    """Backward of the training-mode forward that filled ``ws``: returns (flat gradient buffer, [conv_w grads],
    [bn weight grads], [bn bias grads]) — views into the one flat buffer, in module order."""
    L.require_device()
    grad_out = _cuda_f32(grad_out)
    B, _, H, W_ = depth_shape
    sizes = [t.numel() for t in convs] + [t.numel() for t in bn_w] + [t.numel() for t in bn_b]
    padded = [(n + 63) // 64 * 64 for n in sizes]
    flat = torch.empty(sum(padded), dtype=torch.float32, device=grad_out.device)
    views, o = [], 0
    for n, pn, t in zip(sizes, padded, list(convs) + list(bn_w) + list(bn_b)):
        views.append(flat[o:o + n].view(t.shape))
        o += pn
    n = L.DEPTH_CONVS
    G = L.VetoDepthGrads()
    for i in range(n):
        G.conv_w[i], G.bn_w[i], G.bn_b[i] = views[i].data_ptr(), views[n + i].data_ptr(), views[2 * n + i].data_ptr()
    tensors = [[_cuda_f32(t) for t in group] for group in (convs, bn_w, bn_b)]
    Wst = _depth_weight_struct(tensors[0], tensors[1], tensors[2], bn_mean, bn_var)
    with torch.cuda.device(grad_out.device):
        L.check(L.load().veto_depth_backbone_backward(L.PRECISIONS[L.TRAIN_PRECISION[precision]], ctypes.byref(Wst), grad_out.data_ptr(), B, H, W_,
                                                      ctypes.byref(G), ws.data_ptr(), ws.numel(), L.stream_ptr()),
                "veto_depth_backbone_backward")
    return flat, views[:n], views[n:2 * n], views[2 * n:]


# --------------------------------------------------------------------------------------------
# launch accounting / per-stage device timing (bench.py)
# --------------------------------------------------------------------------------------------
def launch_count() -> int:
    """Kernel launches this thread has enqueued through the library so far."""
    return int(L.load().veto_last_launch_count())


class StageTimer:
    """with StageTimer() as t: ...  -> t.ms[stage], t.launches[stage] (CUDA events around every library launch)."""
    N_TAGS = 24

    def __enter__(self):
        L.check(L.load().veto_profile_begin(L.stream_ptr()), "veto_profile_begin")
        return self

    def __exit__(self, *exc):
        ms = (ctypes.c_double * self.N_TAGS)()
        n = (ctypes.c_int64 * self.N_TAGS)()
        L.check(L.load().veto_profile_end(ms, n), "veto_profile_end")
        names = [L.load().veto_profile_tag_name(i).decode() for i in range(self.N_TAGS)]
        self.ms = {names[i]: float(ms[i]) for i in range(self.N_TAGS) if n[i]}
        self.launches = {names[i]: int(n[i]) for i in range(self.N_TAGS) if n[i]}
        return False


# --------------------------------------------------------------------------------------------
# test hooks
# --------------------------------------------------------------------------------------------
def test_gemm(a, w, bias=None, residual=None, act: int = 0, precision: str = "fp32", scratch=None, c=None):
    """act: low byte activation (0 none, 1 relu, 2 gelu); 0x100 bf16 hi/lo outputs into the scratch; 0x200 reuse the
    operand split already in `scratch` (timing loops)."""
    L.require_device()
    a, w = _cuda_f32(a), _cuda_f32(w)
    M, K = a.shape
    N = w.shape[0]
    if c is None:
        c = torch.empty((M, N), dtype=torch.float32, device=a.device)
    if scratch is None:
        scratch = torch.empty(4 * (M * K + N * K) + 4 * M * N + 256, dtype=torch.uint8, device=a.device)
    L.check(L.load().veto_test_gemm(a.data_ptr(), w.data_ptr(), L.ptr(bias), L.ptr(residual), c.data_ptr(), M, N, K, act,
                                    L.PRECISIONS[precision], scratch.data_ptr(), scratch.numel(), L.stream_ptr()),
            "veto_test_gemm")
    return c


def test_gemm_tn(y, x, precision: str = "bf16x3", split_k: int = 1, geometry=None):
    """out[Nw,Kw] = y[rows,Nw]^T @ x[rows,Kw] through the MN-major tcgen05 weight-gradient kernel."""
    L.require_device()
    y, x = _cuda_f32(y), _cuda_f32(x)
    rows, Nw = y.shape
    Kw = x.shape[1]
    out = torch.empty((Nw, Kw), dtype=torch.float32, device=y.device)
    scratch = torch.empty(4 * (rows * Nw + rows * Kw) + 4 * Nw * Kw * max(1, split_k) + 256, dtype=torch.uint8, device=y.device)
    geo = (ctypes.c_uint32 * 3)(*[int(v) for v in geometry]) if geometry is not None else None
    L.check(L.load().veto_test_gemm_tn(y.data_ptr(), x.data_ptr(), out.data_ptr(), rows, Nw, Kw, L.PRECISIONS[precision],
                                       split_k, geo, scratch.data_ptr(), scratch.numel(), L.stream_ptr()), "veto_test_gemm_tn")
    return out


def test_layernorm(x, w, b):
    L.require_device()
    x = _cuda_f32(x)
    y = torch.empty_like(x)
    L.check(L.load().veto_test_layernorm(x.data_ptr(), _cuda_f32(w).data_ptr(), _cuda_f32(b).data_ptr(), y.data_ptr(),
                                         x.shape[0], L.stream_ptr()), "veto_test_layernorm")
    return y


def test_attention(qkv):
    """qkv [n_seq*19, 1728] -> [n_seq*19, 576]"""
    L.require_device()
    qkv = _cuda_f32(qkv)
    n_seq = qkv.shape[0] // N_TOKENS
    out = torch.empty((qkv.shape[0], T_DIM), dtype=torch.float32, device=qkv.device)
    L.check(L.load().veto_test_attention(qkv.data_ptr(), out.data_ptr(), n_seq, L.stream_ptr()), "veto_test_attention")
    return out


def test_attention_tc(qkv, split: bool = True):
    """tcgen05 attention kernel: qkv [n_seq*19, 1728] -> fp32 [n_seq*19, 576]"""
    L.require_device()
    qkv = _cuda_f32(qkv)
    n_seq = qkv.shape[0] // N_TOKENS
    out = torch.empty((qkv.shape[0], T_DIM), dtype=torch.float32, device=qkv.device)
    scratch = torch.empty(4 * qkv.shape[0] * T_DIM + 256, dtype=torch.uint8, device=qkv.device)
    L.check(L.load().veto_test_attention_tc(qkv.data_ptr(), out.data_ptr(), scratch.data_ptr(), n_seq, int(split),
                                            L.stream_ptr()), "veto_test_attention_tc")
    return out
