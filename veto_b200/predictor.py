"""Drop-ins for the registry-registered ``VETOPredictor`` / ``VETOPredictor_MEET``
(pysgg/modeling/roi_heads/relation_head/roi_relation_predictors.py:3997-4139, 3876-3995, 3661-3874).

Same registration names, constructor ``(config, in_channels)``, ``forward(proposals, rel_pair_idxs,
rel_labels, logger, roi_features=, roi_depth_features=, rel_binarys=)`` signature, 6-tuple return and
state_dict keys as the reference, so reference checkpoints load and ``ROIRelationHead.forward``
(relation_head.py:195-203) can call them unchanged.  The modules hold parameters only; the forward
computation is ``veto_relation_forward`` of libveto_b200.so (sm_100a kernels).  There is no PyTorch or
CPU fallback: without the library or an sm_100 device, forward raises.

``eval()`` forward returns the logits; ``train()`` forward of VETOPredictor returns ``add_losses['rel_loss']``
(roi_relation_predictors.py:4131-4136) as a scalar wired into autograd: ``veto_relation_train_step`` computes the
loss AND every gradient in one library call, and ``loss.backward()`` hands those gradients to the parameters and
to ``roi_depth_features`` (so the depth backbone trains through VETOFeatureExtractor's ROIAlign backward).
VETOPredictor_MEET in ``train()`` mode samples and relabels the pairs of every group head on the device
(``veto_meet_group_labels``) and returns one 'group_k_CE_loss' per head from the same library call.
"""
from __future__ import annotations

import math
import random as _random
from typing import List, Optional

import torch
import torch.nn as nn

from . import config as C
from . import meet_sampling, ops
from .registry import ROI_RELATION_PREDICTOR
from .structures import xyxy_boxes


# ---- hooks a harness may replace (the reference resolves these from the dataset / GloVe files) ----
def get_dataset_statistics(config):
    """obj / rel class name lists (reference: pysgg/data/build.py:27-53).  Uses pysgg when importable,
    otherwise synthesises names from the configured class counts."""
    try:
        from pysgg.data import get_dataset_statistics as ref_stats
    except ImportError:      # the reference is absent (GPU box, tests): errors of an importable reference propagate
        n_obj, n_rel = C.num_classes(config)
        return {"obj_classes": ["__background__"] + [f"obj{i}" for i in range(1, n_obj)],
                "rel_classes": ["__background__"] + [f"rel{i}" for i in range(1, n_rel)]}
    return ref_stats(config)


def obj_edge_vectors(names, wv_dir, wv_dim):
    """GloVe rows for the class names (reference: relation_head/utils_motifs.py:151-171); random when the
    reference / the GloVe files are not available (a checkpoint overwrites them anyway)."""
    try:
        from pysgg.modeling.roi_heads.relation_head.utils_motifs import obj_edge_vectors as ref_vecs
    except ImportError:      # no reference: random rows; a missing GloVe file inside a real install still raises
        return torch.randn(len(names), wv_dim)
    return ref_vecs(names, wv_dir=wv_dir, wv_dim=wv_dim)


REFERENCE_PRED_COUNTS_PATH = "/visinf/home/gsudhakaran/scene_graphs/VETO_rebuttal/pred_counts.pkl"


def load_pred_counts(config):
    """Per-predicate training counts for GLOBAL_SETTING.BETA_LOSS.  The reference reads a hard-coded absolute path
    (roi_relation_predictors.py:4059; the same file ships at the root of its repository); here the path comes from
    the extension key VETO_B200.PRED_COUNTS (a .pkl or .npy), falling back to the reference's path."""
    import os
    import pickle
    import numpy as np
    path = C.get(config, "VETO_B200.PRED_COUNTS", "") or REFERENCE_PRED_COUNTS_PATH
    if not os.path.exists(path):
        raise FileNotFoundError(f"GLOBAL_SETTING.BETA_LOSS needs the predicate counts: set VETO_B200.PRED_COUNTS "
                                f"(tried {path})")
    if path.endswith(".npy"):
        return np.load(path)
    with open(path, "rb") as fin:
        return np.asarray(pickle.load(fin))


def beta_loss_weights(rel_counts, num_rel: int, beta: float = 0.999) -> torch.Tensor:
    """Class-balanced CE weights (roi_relation_predictors.py:4060-4066): counts sorted descending, effective-number
    weighting (1 - beta) / (1 - beta^n) normalised to sum num_rel; computed in the counts' own dtype like the reference."""
    import numpy as np
    counts = np.array(rel_counts, copy=True)
    counts[::-1].sort()
    w = (1.0 - beta) / (1 - (beta ** counts))
    w *= float(num_rel) / np.sum(w)
    return torch.FloatTensor(w)


def xavier_init(m: nn.Linear) -> nn.Linear:
    """pysgg/modeling/utils.py-style xavier_normal_ + zero bias is NOT what the reference uses for rel_out:
    miscellaneous.py:85-93 applies xavier_normal_ to the weight only."""
    nn.init.xavier_normal_(m.weight)
    return m


# ---- parameter containers with the reference's module tree (model_veto.py) ----
class _Attention(nn.Module):
    def __init__(self, dim, heads, dropout):
        super().__init__()
        self.heads = heads
        self.to_qkv = nn.Linear(dim, dim * 3, bias=False)
        self.to_out = nn.Sequential(nn.Linear(dim, dim), nn.Dropout(dropout))


class _FeedForward(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(dim, hidden), nn.GELU(), nn.Dropout(0.0), nn.Linear(hidden, dim),
                                 nn.Dropout(0.0))


class _PreNorm(nn.Module):
    def __init__(self, dim, fn):
        super().__init__()
        self.norm = nn.LayerNorm(dim)
        self.fn = fn


class _PatchEmbed(nn.Module):
    def __init__(self, in_channels, patch):
        super().__init__()
        self.proj_d = nn.Linear(in_channels * 2 * patch ** 2, 512)
        self.proj_v = nn.Linear(in_channels * 2 * patch ** 2, 64)


class _Transformer(nn.Module):
    def __init__(self, config, in_channels):
        super().__init__()
        t = config.MODEL.ROI_RELATION_HEAD.VETOTRANSFORMER
        dim = t.T_INPUT_DIM
        self.patch_embed = _PatchEmbed(in_channels, t.PATCH_SIZE)
        self.cls_token = nn.Parameter(torch.randn(1, 1, dim))
        self.pos_embedding = nn.Parameter(torch.randn(1, 1, dim))
        self.pos_drop = nn.Dropout(t.EMB_DROPOUT)  # model_veto.py:43 (no parameters: state_dict unchanged)
        self.layers = nn.ModuleList([
            nn.ModuleList([_PreNorm(dim, _Attention(dim, t.NHEADS, t.T_DROPOUT)),
                           _PreNorm(dim, _FeedForward(dim, dim * 2))])
            for _ in range(t.ENC_LAYERS)])


class _VETOTransformer(nn.Module):
    def __init__(self, config, in_channels=256):
        super().__init__()
        self.transformer = _Transformer(config, in_channels)


class _Trunk(nn.Module):
    """Everything VETOPredictor and Ensemble share (same parameter names in both)."""

    def _build_trunk(self, config, obj_classes):
        t = config.MODEL.ROI_RELATION_HEAD.VETOTRANSFORMER
        dim = t.T_INPUT_DIM
        if (dim, t.NHEADS, t.PATCH_SIZE, config.MODEL.ROI_RELATION_HEAD.POOLER_RESOLUTION) != (576, 6, 2, 8):
            raise RuntimeError("veto_b200 kernels are built for T_INPUT_DIM 576 / NHEADS 6 / PATCH_SIZE 2 / "
                               "POOLER_RESOLUTION 8 (configs/VETO_final.yaml)")
        self.embed_dim = 200
        self.obj_embed = nn.Embedding(len(obj_classes), self.embed_dim)
        self.class_projection = nn.Sequential(nn.Linear(400, dim), nn.ReLU(inplace=True))
        with torch.no_grad():
            self.obj_embed.weight.copy_(obj_edge_vectors(obj_classes, config.GLOVE_DIR, self.embed_dim))
        self.bbox_embed = nn.Sequential(nn.Linear(9, 32), nn.ReLU(inplace=True), nn.Dropout(0.1),
                                        nn.Linear(32, 128), nn.ReLU(inplace=True), nn.Dropout(0.1))
        self.pos_embed = nn.Sequential(nn.BatchNorm1d(4, momentum=0.001), nn.Linear(4, 128), nn.ReLU(inplace=True),
                                       nn.Dropout(0.1))
        self.location_projection = nn.Sequential(nn.Linear(256, dim), nn.ReLU(inplace=True))
        self.fusion_transformer = _VETOTransformer(config, in_channels=256)
        self.n_layers = t.ENC_LAYERS
        self.precision = C.get(config, "VETO_B200.PRECISION", "f16c8")
        self.chunk_pairs = int(C.get(config, "VETO_B200.CHUNK_PAIRS", 0) or 0)
        self._packed = None
        self._packed_key = None

    def _trunk_tensors(self):
        return {k: v for k, v in self.state_dict(keep_vars=True).items()}

    def _pack(self, rel_w: torch.Tensor, rel_b: torch.Tensor, training: bool = False) -> ops.PackedWeights:
        from .lib import TRAIN_PRECISION
        precision = TRAIN_PRECISION[self.precision] if training else self.precision
        tensors = self._trunk_tensors()
        key = (precision, tuple((k, t.data_ptr(), t._version) for k, t in tensors.items()),
               rel_w.data_ptr(), rel_w._version, rel_b._version)
        if self._packed is None or self._packed_key != key:
            cfg = ops.make_config(self.num_obj_cls, rel_w.shape[0], precision, layers=self.n_layers)
            self._packed = ops.PackedWeights(cfg, tensors, rel_w, rel_b)
            self._packed_key = key
        return self._packed

    def _relation_logits(self, proposals, rel_pair_idxs, roi_features, roi_depth_features, rel_w, rel_b,
                         labels=None, obj_logits=None, freq_bias=None):
        n_boxes = [len(p) for p in proposals]
        boxes = torch.cat([xyxy_boxes(p) for p in proposals], 0)
        subj, obj = ops.globalize_pairs(rel_pair_idxs, n_boxes)
        pw = self._pack(rel_w, rel_b)
        return ops.relation_forward(pw, boxes, roi_features, roi_depth_features, subj, obj, labels=labels,
                                    obj_logits=obj_logits, freq_bias=freq_bias, chunk_pairs=self.chunk_pairs)


    # ---- training branch (roi_relation_predictors.py:4131-4136 / 3812-3846) ----
    def _trained_params(self):
        """(state_dict key, parameter) of everything the losses depend on, in the order of ops.grad_fields."""
        named = dict(self.named_parameters())
        return [(key, named[key]) for _, _, key in ops.grad_fields(self.n_layers)]

    def _train_forward(self, proposals, rel_pair_idxs, roi_features, roi_depth_features, hard, soft, heads,
                       rel_labels=None, class_weight=None, head_labels=None):
        """One veto_relation_train_step wired into autograd.  `heads`: the classifier Linear modules whose rows,
        concatenated, are the library's rel_out ([rel_out] for VETOPredictor, the group heads for MEET); with more
        than one head `head_labels` [n_heads, R] (-1 = pair not in that head's loss) replaces `rel_labels` and the
        result is the vector of head losses."""
        n_boxes = [len(p) for p in proposals]
        rel_counts = [int(r.shape[0]) for r in rel_pair_idxs]
        boxes = torch.cat([xyxy_boxes(p) for p in proposals], 0)
        subj, obj = ops.globalize_pairs(rel_pair_idxs, n_boxes)
        keyed = self._trained_params()
        head_sizes = [int(m.weight.shape[0]) for m in heads]
        n_out = sum(head_sizes)
        params = [p for _, p in keyed] + [m.weight for m in heads] + [m.bias for m in heads]
        tr = self.fusion_transformer.transformer
        p_attn = {float(layer[0].fn.to_out[1].p) for layer in tr.layers}
        if len(p_attn) != 1:
            raise RuntimeError("veto_b200: every Attention.to_out Dropout must use the same p")
        seed = int(torch.randint(0, 2 ** 62, (1,)).item())  # CPU generator: follows torch.manual_seed, no device sync
        bn = self.pos_embed[0]
        dim = heads[0].weight.shape[1]

        def run(want_depth, want_rgb):
            if len(heads) == 1:
                rel_w, rel_b = heads[0].weight, heads[0].bias
            else:
                rel_w = torch.cat([m.weight.detach() for m in heads], 0)
                rel_b = torch.cat([m.bias.detach() for m in heads], 0)
            pw = self._pack(rel_w, rel_b, training=True)
            n_trunk = sum(p.numel() for _, p in keyed)
            flat = torch.empty(n_trunk + n_out * dim + n_out, dtype=torch.float32, device=boxes.device)
            views, off = [], 0
            for _, p in keyed:
                views.append(flat[off:off + p.numel()].view(p.shape))
                off += p.numel()
            g_w = flat[off:off + n_out * dim].view(n_out, dim)    # gradient of the concatenated classifier
            g_b = flat[off + n_out * dim:]
            row = 0
            for n in head_sizes:                                   # a head's rows are a contiguous slice of it
                views.append(g_w[row:row + n])
                row += n
            row = 0
            for n in head_sizes:
                views.append(g_b[row:row + n])
                row += n
            grads = {key: v for (key, _), v in zip(keyed, views)}
            loss, g_depth, g_rgb, _ = ops.relation_train_step(
                pw, boxes, roi_features.detach(), roi_depth_features.detach(), subj, obj, rel_labels, rel_counts, n_boxes,
                grads, g_w, g_b, labels=hard, obj_logits=soft, class_weight=class_weight,
                p_pos=float(self.pos_embed[3].p), p_emb=float(tr.pos_drop.p), p_attn=next(iter(p_attn)), seed=seed,
                bn_momentum=float(bn.momentum), bn_running_mean=bn.running_mean if bn.track_running_stats else None,
                bn_running_var=bn.running_var if bn.track_running_stats else None, want_roi_depth_grad=want_depth,
                want_roi_rgb_grad=want_rgb, head_sizes=head_sizes if len(heads) > 1 else None, head_labels=head_labels)
            if bn.track_running_stats:
                bn.num_batches_tracked += 1
            return loss, g_depth, g_rgb, flat, views

        return _TrainStep.apply(self, run, roi_features, roi_depth_features, *params)


class _TrainStep(torch.autograd.Function):
    """The training losses with all of their gradients computed eagerly by veto_relation_train_step; backward() only
    scales them by the incoming gradient and hands them to autograd (parameters are inputs, so DDP / optimizers see
    .grad).  Vanilla: one scalar rel_loss.  MEET: a vector of group losses whose gradients the library computed for
    their SUM — what the reference's trainer back-propagates (tools/relation_train_net.py:451: losses = sum(...));
    unequal incoming weights cannot be honoured and poison the gradients with NaN instead of being silently wrong."""

    @staticmethod
    def forward(ctx, module, run, roi_rgb, roi_depth, *params):
        loss, g_depth, g_rgb, flat, grads = run(want_depth=ctx.needs_input_grad[3], want_rgb=ctx.needs_input_grad[2])
        ctx.flat, ctx.grads, ctx.g_depth, ctx.g_rgb, ctx.module = flat, grads, g_depth, g_rgb, module
        return loss.reshape(()) if loss.numel() == 1 else loss

    @staticmethod
    def backward(ctx, gout):
        if gout.dim():
            uniform = (gout == gout[0]).all()
            gout = torch.where(uniform, gout[0], torch.full_like(gout[0], float("nan")))
        ctx.flat.mul_(gout)  # one kernel over the flat gradient buffer; `grads` are views of it
        # data-parallel hook (veto_b200.distributed.allreduce_flat): every gradient of the head is final here, before
        # autograd walks on into the ROIAlign and depth-backbone backward — the exchange can run underneath them
        hook = getattr(ctx.module, "grad_sync", None)
        if hook is not None:
            hook(ctx.flat)
        scale = lambda t: None if t is None else t * gout
        # hand the views over without keeping a reference: autograd's AccumulateGrad then adopts them as .grad instead
        # of cloning each one (about 90 small copy kernels per step), and every .grad stays a view of the one flat
        # buffer that the all-reduce / clipping / optimizer kernels sweep
        grads, ctx.grads, ctx.flat = ctx.grads, None, None
        return (None, None, scale(ctx.g_rgb), scale(ctx.g_depth)) + tuple(grads)


def _mode(config) -> str:
    if config.MODEL.ROI_RELATION_HEAD.USE_GT_BOX:
        return "predcls" if config.MODEL.ROI_RELATION_HEAD.USE_GT_OBJECT_LABEL else "sgcls"
    return "sgdet"


@ROI_RELATION_PREDICTOR.register("VETOPredictor")
class VETOPredictor(_Trunk):
    """roi_relation_predictors.py:3997-4139."""

    def __init__(self, config, in_channels):
        super().__init__()
        self.mode = _mode(config)
        statistics = get_dataset_statistics(config)
        self.obj_classes, self.rel_classes = statistics["obj_classes"], statistics["rel_classes"]
        self.num_obj_cls, self.num_rel_cls = len(self.obj_classes), len(self.rel_classes)
        self.obj_embed2 = nn.Embedding(self.num_obj_cls, 200)  # unused by forward, kept for checkpoint parity
        self._build_trunk(config, self.obj_classes)
        dim = config.MODEL.ROI_RELATION_HEAD.VETOTRANSFORMER.T_INPUT_DIM
        self.rel_out = xavier_init(nn.Linear(dim, self.num_rel_cls, bias=True))
        self.beta_loss = bool(config.GLOBAL_SETTING.BETA_LOSS)
        if self.beta_loss:
            rel_class_weights = beta_loss_weights(load_pred_counts(config), self.num_rel_cls)
        else:
            rel_class_weights = torch.ones(self.num_rel_cls)
        self.criterion_loss_rel = nn.CrossEntropyLoss(weight=rel_class_weights)
        self.criterion_loss = nn.CrossEntropyLoss()
        self.use_freq_bias = bool(C.get(config, "VETO_B200.FREQ_BIAS", False))
        self.freq_bias_table = None  # [num_obj^2, num_rel] fp32, set by the caller when use_freq_bias

    def forward(self, proposals, rel_pair_idxs, rel_labels, logger, roi_features=None, roi_depth_features=None,
                rel_binarys=None):
        if self.training:
            if self.mode == "predcls":
                hard, soft = torch.cat([p.get_field("labels") for p in proposals], 0).long(), None
            else:
                hard, soft = None, torch.cat([p.get_field("predict_logits") for p in proposals], 0).detach()
            add_losses = {}
            if self.mode != "predcls":  # :4131-4133: CE of the (detached) one-hot predictions, a constant
                pred = torch.cat([p.get_field("pred_labels") for p in proposals], 0).detach().long()
                fg = torch.cat([p.get_field("labels") for p in proposals], 0).long()
                add_losses["obj_loss"] = self.criterion_loss(nn.functional.one_hot(pred, self.num_obj_cls).float(), fg)
            add_losses["rel_loss"] = self._train_forward(
                proposals, rel_pair_idxs, roi_features, roi_depth_features, hard, soft, [self.rel_out],
                rel_labels=torch.cat(list(rel_labels), 0).long(), class_weight=self.criterion_loss_rel.weight)
            return None, None, add_losses, None, None, None
        if self.mode == "predcls":
            obj_labels = torch.cat([p.get_field("labels") for p in proposals], 0).long()
            hard, soft = obj_labels, None
        else:
            soft = torch.cat([p.get_field("predict_logits") for p in proposals], 0).detach()
            obj_labels = torch.cat([p.get_field("pred_labels") for p in proposals], 0).detach().long()
            hard = None
        obj_dists = nn.functional.one_hot(obj_labels, self.num_obj_cls).float()
        fb = None
        if self.use_freq_bias:
            if self.freq_bias_table is None or hard is None:
                raise RuntimeError("FREQ_BIAS needs freq_bias_table and hard labels (predcls)")
            fb = self.freq_bias_table
        rel_dists = self._relation_logits(proposals, rel_pair_idxs, roi_features, roi_depth_features,
                                          self.rel_out.weight, self.rel_out.bias, labels=hard, obj_logits=soft,
                                          freq_bias=fb)
        obj_dists = obj_dists.split([len(p) for p in proposals], dim=0)
        rel_dists = rel_dists.split([len(r) for r in rel_pair_idxs], dim=0)
        return obj_dists, rel_dists, {}, None, None, None


def incre_idx_list(group_sizes: List[int], num_rel: int) -> List[int]:
    """SHA_GCL_extra/extra_function_utils.py:39-52: predicate id -> 1-based group id (0 = background)."""
    out = [0] * num_rel
    c = 1
    for g, n in enumerate(group_sizes):
        for _ in range(n):
            out[c] = g + 1
            c += 1
    return out


class Ensemble(_Trunk):
    """roi_relation_predictors.py:3661-3874 (ensemble_type 'group')."""

    def __init__(self, config, mode, params, group_num, exp_per_group, group_element_number_list, idx_list):
        super().__init__()
        self.mode = mode
        self.num_obj_cls = len(params["obj_classes"])
        self.num_rel_cls = len(params["rel_classes"])
        self.incre_idx_list = idx_list
        self._build_trunk(config, params["obj_classes"])
        dim = config.MODEL.ROI_RELATION_HEAD.VETOTRANSFORMER.T_INPUT_DIM
        self.group_num = group_num
        self.experts_per_group = exp_per_group
        self.expert_group = bool(config.ENSEMBLE_LEARNING.EXPERT_GROUP)
        self.group_outs = [n + 2 for n in group_element_number_list]
        self.rel_out = nn.ModuleList([])
        self.rel_out_group = nn.ModuleList([])
        if self.expert_group:
            for _ in range(exp_per_group):
                self.rel_out = nn.ModuleList([xavier_init(nn.Linear(dim, n, bias=True)) for n in self.group_outs])
                self.rel_out_group.append(self.rel_out)
        else:
            for n in self.group_outs:
                self.rel_out.append(xavier_init(nn.Linear(dim, n, bias=True)))
        self.CE_loss = nn.CrossEntropyLoss()
        self.criterion_loss = nn.CrossEntropyLoss()
        self.nms_thresh = config.TEST.RELATION.LATER_NMS_PREDICTION_THRES

    def _forward_train(self, proposals, rel_pair_idxs, rel_labels, roi_features, roi_depth_features, obj_preds,
                       cur_chosen_matrix):
        """roi_relation_predictors.py:3806-3848 in train() mode: per-group relabelling of the sampled pairs and one
        CE per head ('group_k_CE_loss'; with EXPERT_GROUP every expert j of group k sees group k's pairs, :3834-3840).
        `rel_labels` is the concatenated label tensor, `cur_chosen_matrix` = expert_dist of VETOPredictor_MEET."""
        if cur_chosen_matrix is None:
            raise RuntimeError("Ensemble in train() mode needs cur_chosen_matrix (VETOPredictor_MEET.forward builds it)")
        dev = roi_features.device
        if isinstance(cur_chosen_matrix, meet_sampling.LazyExpertDist):
            table = cur_chosen_matrix.table                       # [G, R] head-local labels, built on the device
        else:   # chosen-rows lists as the reference passes them (a caller that sampled on the host): relabel on the host
            labels_host = rel_labels if isinstance(rel_labels, (list, tuple)) else rel_labels.tolist()
            table = torch.from_numpy(meet_sampling.group_local_labels(labels_host, cur_chosen_matrix[0],
                                                                      self.incre_idx_list)).to(dev, non_blocking=True)
        heads = self._head_sets()
        # group index of every head, in _head_sets order (expert-major)
        per_head = [k for _ in range(self.experts_per_group if self.expert_group else 1) for k in range(self.group_num)]
        head_labels = table if per_head == list(range(self.group_num)) else table[per_head]
        add_losses = {}
        if self.mode != "predcls":  # :3826-3830: CE of the detached detector logits, a constant of the step
            obj_logits = torch.cat([p.get_field("predict_logits") for p in proposals], 0).detach()
            fg = torch.cat([p.get_field("labels") for p in proposals], 0).long()
            add_losses["obj_loss"] = self.criterion_loss(obj_logits, fg)
        losses = self._train_forward(proposals, rel_pair_idxs, roi_features, roi_depth_features, obj_preds, None,
                                     [m for _, m in heads], head_labels=head_labels)
        for i, (name, _) in enumerate(heads):
            add_losses[name + "_CE_loss"] = losses[i]
        return None, None, add_losses, None

    def _head_sets(self):
        if self.expert_group:
            return [(f"group_%d%d" % (k, j + 1), self.rel_out_group[j][k]) for j in range(self.experts_per_group)
                    for k in range(self.group_num)]
        return [("group_%d" % k, self.rel_out[k]) for k in range(self.group_num)]

    def forward(self, proposals, rel_pair_idxs, rel_labels, logger, roi_features=None, roi_depth_features=None,
                cur_chosen_matrix=None):
        soft = None
        if self.mode == "predcls":
            obj_preds = torch.cat([p.get_field("labels") for p in proposals], 0).long()
            obj_dists = nn.functional.one_hot(obj_preds, self.num_obj_cls).float()
        else:
            obj_labels = torch.cat([p.get_field("pred_labels") for p in proposals], 0).detach().long()
            obj_dists = nn.functional.one_hot(obj_labels, self.num_obj_cls).float()
            if self.mode == "sgdet" and not self.training:  # use_decoder_nms (:3776-3781)
                boxes_per_cls = torch.cat([p.get_field("boxes_per_cls") for p in proposals], 0)
                obj_preds = ops.obj_nms_per_cls(torch.softmax(obj_dists, -1), boxes_per_cls,
                                                [len(p) for p in proposals], self.nms_thresh)
            else:
                obj_preds = obj_dists[:, 1:].max(1)[1] + 1  # :3783
        if self.training:
            return self._forward_train(proposals, rel_pair_idxs, rel_labels, roi_features, roi_depth_features,
                                       obj_preds, cur_chosen_matrix)
        heads = self._head_sets()
        # all expert heads as ONE [sum(n_k+2), 576] classifier GEMM, split afterwards
        w = torch.cat([m.weight for _, m in heads], 0)
        b = torch.cat([m.bias for _, m in heads], 0)
        key = tuple((m.weight.data_ptr(), m.weight._version, m.bias._version) for _, m in heads)
        if getattr(self, "_cat_key", None) != key:
            self._cat_w, self._cat_b, self._cat_key = w.detach().contiguous(), b.detach().contiguous(), key
        logits = self._relation_logits(proposals, rel_pair_idxs, roi_features, roi_depth_features, self._cat_w,
                                       self._cat_b, labels=obj_preds)
        rel_dists, off = {}, 0
        for name, m in heads:
            n = m.weight.shape[0]
            rel_dists[name] = logits[:, off:off + n]
            off += n
        obj_dists = obj_dists.split([len(p) for p in proposals], dim=0)
        return obj_dists, rel_dists, {}, None


@ROI_RELATION_PREDICTOR.register("VETOPredictor_MEET")
class VETOPredictor_MEET(nn.Module):
    """roi_relation_predictors.py:3876-3995."""

    def __init__(self, config, in_channels):
        super().__init__()
        self.mode = _mode(config)
        stats = get_dataset_statistics(config)
        self.params = {"statistics": stats, "obj_classes": stats["obj_classes"], "rel_classes": stats["rel_classes"]}
        ds = config.GLOBAL_SETTING.DATASET_CHOICE
        split = config.GCL_SETTING.GROUP_SPLIT_MODE
        if (ds, split) not in C.GROUP_SPLITS:
            raise KeyError(f"unknown group split {(ds, split)}")
        self.max_group_element_number_list = list(C.GROUP_SPLITS[(ds, split)])
        self.incre_idx_list = incre_idx_list(self.max_group_element_number_list, len(stats["rel_classes"]))
        self.num_groups = len(self.max_group_element_number_list)
        self.experts_per_group = 3 if config.ENSEMBLE_LEARNING.EXPERT_GROUP else 1
        self.ensemble_type = config.ENSEMBLE_LEARNING.TYPE
        self.zero_label_padding_mode = config.GCL_SETTING.ZERO_LABEL_PADDING_MODE
        self.sample_rate_matrix = meet_sampling.sample_rate_matrix(ds, self.max_group_element_number_list)
        self._tables = None
        self.reference_draws = bool(C.get(config, "VETO_B200.MEET_REFERENCE_DRAWS", False))
        self.model = Ensemble(config, self.mode, self.params, self.num_groups, self.experts_per_group,
                              self.max_group_element_number_list, self.incre_idx_list)

    def _sampling_tables(self, device):
        """(incre_idx int32 [num_rel], rates float64 [G, num_rel], local_label int32 [G, num_rel]) on `device`."""
        if self._tables is None or self._tables[0].device != device:
            import numpy as np
            self._tables = (torch.tensor(self.incre_idx_list, dtype=torch.int32, device=device),
                            torch.from_numpy(np.ascontiguousarray(self.sample_rate_matrix, dtype=np.float64)).to(device),
                            torch.from_numpy(meet_sampling.local_label_table(self.incre_idx_list, self.num_groups)).to(device))
        return self._tables

    def forward(self, proposals, rel_pair_idxs, rel_labels, logger, roi_features=None, roi_depth_features=None,
                rel_binarys=None):
        if self.training:
            # :3926-3969 — the reference walks the pairs in Python with one .item() sync each; here one kernel draws and
            # relabels every pair on the device.  The seed comes from Python's `random` (the reference's source of
            # randomness for this step), so random.seed(...) makes a run reproducible; no device sync.
            labels = torch.cat(list(rel_labels), 0).long()
            tables = self._sampling_tables(labels.device)
            draws = heads = None
            if self.reference_draws:
                # parity mode (VETO_B200.MEET_REFERENCE_DRAWS): replay the reference's own draws from `random`, in its
                # order, so that a run seeded like the reference samples the same pairs (one D2H of the labels)
                draws, heads = meet_sampling.reference_draws(labels.tolist(), self.num_groups, self.zero_label_padding_mode)
                draws, heads = torch.from_numpy(draws), torch.from_numpy(heads)
            table = ops.meet_group_labels(labels, *tables, self.zero_label_padding_mode, seed=_random.getrandbits(63),
                                          draws=draws, bg_heads=heads)
            expert_dist = meet_sampling.LazyExpertDist(table)
            _, _, add_losses, _ = self.model(proposals, rel_pair_idxs, labels, logger, roi_features=roi_features,
                                             roi_depth_features=roi_depth_features, cur_chosen_matrix=expert_dist)
            return None, None, dict(add_losses), self.incre_idx_list, expert_dist, None
        obj_dists, rel_dists, add_losses, _ = self.model(proposals, rel_pair_idxs, rel_labels, logger,
                                                         roi_features=roi_features,
                                                         roi_depth_features=roi_depth_features)
        return obj_dists, dict(rel_dists), dict(add_losses), self.incre_idx_list, None, {}
