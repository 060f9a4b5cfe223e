"""Minimal BoxList: the input container of the drop-in boundary.

When the reference's ``pysgg`` is importable the drop-in modules receive its own
``pysgg.structures.bounding_box.BoxList`` (duck-typed: ``bbox``, ``size``, ``mode``, ``get_field``,
``add_field``, ``convert``, ``__len__``); this class offers the same surface for use without the
reference (GPU box, tests, bench).  Semantics follow pysgg/structures/bounding_box.py:9-259
(the +1 ``TO_REMOVE`` convention of ``convert('xywh')`` :72-75 and ``area()`` :249-259).
"""
from __future__ import annotations

import torch


class BoxList:
    def __init__(self, bbox, image_size, mode: str = "xyxy"):
        device = bbox.device if isinstance(bbox, torch.Tensor) else torch.device("cpu")
        bbox = torch.as_tensor(bbox, dtype=torch.float32, device=device)
        if bbox.ndimension() != 2 or bbox.size(-1) != 4:
            raise ValueError(f"bbox should be [N,4], got {tuple(bbox.shape)}")
        if mode not in ("xyxy", "xywh"):
            raise ValueError("mode should be 'xyxy' or 'xywh'")
        self.bbox = bbox
        self.size = image_size  # (image_width, image_height)
        self.mode = mode
        self.extra_fields = {}

    def add_field(self, field, field_data):
        self.extra_fields[field] = field_data

    def get_field(self, field):
        return self.extra_fields[field]

    def has_field(self, field):
        return field in self.extra_fields

    def fields(self):
        return list(self.extra_fields.keys())

    def convert(self, mode):
        if mode == self.mode:
            return self
        b = self.bbox
        if mode == "xywh":  # from xyxy
            out = torch.stack([b[:, 0], b[:, 1], b[:, 2] - b[:, 0] + 1, b[:, 3] - b[:, 1] + 1], 1)
        else:  # xywh -> xyxy
            out = torch.stack([b[:, 0], b[:, 1], b[:, 0] + (b[:, 2] - 1).clamp(min=0),
                               b[:, 1] + (b[:, 3] - 1).clamp(min=0)], 1)
        r = BoxList(out, self.size, mode)
        r.extra_fields = dict(self.extra_fields)
        return r

    def area(self):
        b = self.bbox
        if self.mode == "xyxy":
            return (b[:, 2] - b[:, 0] + 1) * (b[:, 3] - b[:, 1] + 1)
        return b[:, 2] * b[:, 3]

    def to(self, device):
        r = BoxList(self.bbox.to(device), self.size, self.mode)
        for k, v in self.extra_fields.items():
            r.add_field(k, v.to(device) if hasattr(v, "to") else v)
        return r

    def __len__(self):
        return self.bbox.shape[0]

    def __repr__(self):
        return f"BoxList(num_boxes={len(self)}, image_width={self.size[0]}, image_height={self.size[1]}, mode={self.mode})"


def xyxy_boxes(proposal) -> torch.Tensor:
    """[N,4] xyxy boxes of a (reference or local) BoxList."""
    if proposal.mode == "xyxy":
        return proposal.bbox
    return proposal.convert("xyxy").bbox
