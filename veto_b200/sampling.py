"""Drop-in for ``RelationSampling.prepare_test_pairs``
(pysgg/modeling/roi_heads/relation_head/sampling.py:31-52): candidate-pair enumeration on the device in one
launch for the whole batch, instead of a per-image loop of ones/eye/nonzero/sort.

Tie-breaking over the MAX_PROPOSAL_PAIR cap: the reference uses an unstable torch.sort on
pred_scores[i]*pred_scores[j]; every pair ties with its mirror.  Here the order is defined as
(product descending, row-major pair index ascending) — what the reference's CPU path produces.
"""
from __future__ import annotations

import torch

from . import config as C
from . import ops
from .structures import xyxy_boxes


class PendingRelSample:
    """A gtbox_relsample in flight (RelationSampling.gtbox_relsample_async)."""

    def __init__(self, proposals, pairs, labels, counts_host, binaries, event, batch_size):
        self.proposals, self.pairs, self.labels, self.counts_host = proposals, pairs, labels, counts_host
        self.binaries, self.event, self.batch_size = binaries, event, batch_size

    def result(self):
        """(proposals, rel_labels, rel_idx_pairs, rel_sym_binarys), the return value of gtbox_relsample."""
        self.event.synchronize()
        totals = self.counts_host[:, 1].tolist()
        bs = self.batch_size
        rel_idx_pairs = [self.pairs[b * bs: b * bs + n] for b, n in enumerate(totals)]
        rel_labels = [self.labels[b * bs: b * bs + n] for b, n in enumerate(totals)]
        return self.proposals, rel_labels, rel_idx_pairs, self.binaries


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

    def gtbox_relsample_async(self, proposals, targets) -> "PendingRelSample":
        """Enqueue gtbox_relsample for a batch and return without waiting: the one thing the host needs (the row count
        per image) travels to pinned memory behind the kernel, guarded by an event.  A training loop issues the
        sampling of step k+1 right at the start of step k — it depends on the targets only — and calls ``result()``
        one step later, when the event has long fired: the sampler's host sync then never drains the GPU queue."""
        assert self.use_gt_box
        num_pos = int(self.batch_size_per_image * self.positive_fraction)
        for p, t in zip(proposals, targets):
            assert p.bbox.shape[0] == t.bbox.shape[0]
            p.add_field("locating_match", torch.ones(len(p), device=p.bbox.device))          # :73-75
        seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        pairs, labels, counts, binaries = ops.relsample_gtbox([t.get_field("relation") for t in targets],
                                                              self.batch_size_per_image, num_pos, seed)
        counts_host = torch.empty(counts.shape, dtype=counts.dtype).pin_memory()
        counts_host.copy_(counts, non_blocking=True)
        event = torch.cuda.Event()
        event.record()
        return PendingRelSample(proposals, pairs, labels, counts_host, binaries, event, self.batch_size_per_image)

    def gtbox_relsample(self, proposals, targets):
        """sampling.py:54-107: (proposals, rel_labels, rel_idx_pairs, rel_sym_binarys) for training on ground-truth
        boxes — one launch for the batch and ONE host sync (the row counts), instead of a per-image loop with three
        nonzero syncs and two randperms.  The random subset / order comes from a counter-based hash seeded from torch's
        CPU generator (torch.manual_seed reproduces a run); foreground rows keep the reference's row-major order when
        all of them fit."""
        return self.gtbox_relsample_async(proposals, targets).result()

    def detect_relsample(self, proposals, targets):
        """sampling.py:109-309 (with motif_rel_fg_bg_sampling): (proposals, rel_labels, rel_labels_all, rel_idx_pairs,
        rel_sym_binarys) for training on DETECTED boxes — IoU matching against the ground truth, per-relation
        candidate pairs, the IoU-weighted draw, the quality-ranked background pool — one launch and one host sync per
        batch instead of a Python loop over images and ground-truth relations."""
        num_pos = int(self.batch_size_per_image * self.positive_fraction)
        self.num_pos_per_img = num_pos
        seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        trip, corr, counts, binaries, locs = ops.relsample_detect(
            [xyxy_boxes(p) for p in proposals], [p.get_field("labels") for p in proposals],
            [p.get_field("pred_scores") for p in proposals], [xyxy_boxes(t) for t in targets],
            [t.get_field("labels") for t in targets], [t.get_field("relation") for t in targets], self.fg_thres,
            self.require_overlap and not self.use_gt_box, self.num_sample_per_gt_rel, self.batch_size_per_image, num_pos, seed)
        totals = counts[:, 1].tolist()
        bs = self.batch_size_per_image
        rel_idx_pairs, rel_labels, rel_labels_all = [], [], []
        for b, (p, t, n) in enumerate(zip(proposals, targets, totals)):
            p.add_field("locating_match", locs[b])                                          # :137-141
            rows = trip[b * bs: b * bs + n]
            rel_idx_pairs.append(rows[:, :2])
            rel_labels.append(rows[:, 2])
            if t.has_field("relation_non_masked"):                                          # :160-169
                rel_map = t.get_field("relation_non_masked")
                gt_rel_idx = torch.nonzero(rel_map != 0)
                c = corr[b * bs: b * bs + n]
                fg = gt_rel_idx[c[c >= 0]]
                fg_labels = rel_map[fg[:, 0], fg[:, 1]].long()
                rel_labels_all.append(torch.cat((fg_labels, torch.zeros(int((c < 0).sum()), dtype=torch.long, device=c.device))))
        if not rel_labels_all:
            rel_labels_all = rel_labels
        return proposals, rel_labels, rel_labels_all, rel_idx_pairs, binaries


def make_roi_relation_samp_processor(cfg):
    """sampling.py:312-324."""
    rh = cfg.MODEL.ROI_RELATION_HEAD
    return RelationSampling(
        fg_thres=C.get(cfg, "MODEL.ROI_HEADS.FG_IOU_THRESHOLD", 0.5),
        require_overlap=C.get(cfg, "MODEL.ROI_RELATION_HEAD.REQUIRE_BOX_OVERLAP", False),
        num_sample_per_gt_rel=C.get(cfg, "MODEL.ROI_RELATION_HEAD.NUM_SAMPLE_PER_GT_REL", 4),
        batch_size_per_image=C.get(cfg, "MODEL.ROI_RELATION_HEAD.BATCH_SIZE_PER_IMAGE", 1024),
        positive_fraction=C.get(cfg, "MODEL.ROI_RELATION_HEAD.POSITIVE_FRACTION", 0.25),
        max_proposal_pairs=rh.MAX_PROPOSAL_PAIR,
        use_gt_box=rh.USE_GT_BOX,
        test_overlap=cfg.TEST.RELATION.REQUIRE_OVERLAP,
    )
