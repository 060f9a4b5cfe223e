"""Torch-side builders of the synthetic workloads shared by bench.py, tools/ and the tests: the config
node, device BoxLists with the fields ROIRelationHead sets (relation_head.py:104-111), device feature maps and a
predictor loaded from a seeded state.  Numpy generators live in synth.py; nothing here touches oracle/ or tests/."""
from __future__ import annotations

import torch

from . import config as vcfg
from . import registry, synth
from .structures import BoxList


def onehot_logits(labels: torch.Tensor, num_obj: int, fill: float = 1000.0) -> torch.Tensor:
    """to_onehot (model_kern.py:266-281) as used by relation_head.py:104-111: -fill everywhere, +fill at the label."""
    out = torch.full((labels.shape[0], num_obj), -fill, dtype=torch.float32, device=labels.device)
    out[torch.arange(labels.shape[0], device=labels.device), labels] = fill
    return out


def make_cfg(predictor="VETOPredictor", mode="predcls", dataset="VG", max_pairs=2048, require_overlap=False,
             precision="fp32", chunk_pairs=0):
    cfg = vcfg.default_cfg()
    cfg.merge_from_list([
        "MODEL.ROI_RELATION_HEAD.PREDICTOR", predictor,
        "MODEL.ROI_RELATION_HEAD.USE_GT_BOX", mode in ("predcls", "sgcls"),
        "MODEL.ROI_RELATION_HEAD.USE_GT_OBJECT_LABEL", mode == "predcls",
        "MODEL.ROI_RELATION_HEAD.MAX_PROPOSAL_PAIR", max_pairs,
        "TEST.RELATION.REQUIRE_OVERLAP", require_overlap,
        "GLOBAL_SETTING.DATASET_CHOICE", dataset,
        "ENSEMBLE_LEARNING.ENABLED", predictor.endswith("MEET"),
        "VETO_B200.PRECISION", precision,
        "VETO_B200.CHUNK_PAIRS", chunk_pairs,
    ])
    return cfg


def boxlists(batch, device, num_obj):
    out = []
    for i in range(batch["B"]):
        bl = BoxList(torch.from_numpy(batch["boxes"][i]).to(device), (batch["W"], batch["H"]), mode="xyxy")
        lab = torch.from_numpy(batch["labels"][i]).to(device)
        bl.add_field("labels", lab)
        if batch["mode"] == "predcls":
            bl.add_field("predict_logits", onehot_logits(lab, num_obj))
            bl.add_field("pred_scores", torch.ones(len(lab), device=device))
            bl.add_field("pred_labels", lab)
        else:
            bl.add_field("predict_logits", torch.from_numpy(batch["predict_logits"][i]).to(device))
            bl.add_field("pred_scores", torch.from_numpy(batch["pred_scores"][i]).to(device))
            bl.add_field("pred_labels", torch.from_numpy(batch["pred_labels"][i]).to(device))
            if "boxes_per_cls" in batch:
                bl.add_field("boxes_per_cls", torch.from_numpy(batch["boxes_per_cls"][i]).to(device))
        out.append(bl)
    return out


def device_features(batch, device):
    feats = [torch.from_numpy(f).to(device) for f in batch["feats"]]
    feats.append(torch.zeros(batch["B"], feats[0].shape[1], 1, 1, device=device))  # P6: present, unused
    return feats, torch.from_numpy(batch["depth"]).to(device)


def random_features(n_images, H, W, device, generator, channels=synth.CHANNELS):
    """Device-generated N(0,1) FPN maps P2..P5 (+ the unused P6 placeholder) and a relu(N(0,1)) depth map for the
    throughput legs (the numpy generator of synth.make_batch is too slow for 32-image batches)."""
    rgb_hw, depth_hw = synth.fpn_shapes(H, W)
    feats = [torch.randn((n_images, channels) + hw, generator=generator, device=device) for hw in rgb_hw]
    feats.append(torch.zeros(n_images, channels, 1, 1, device=device))
    depth = torch.relu(torch.randn((n_images, channels) + depth_hw, generator=generator, device=device))
    return feats, depth


def build_predictor(cfg, state_np, device):
    pred = registry.make_roi_relation_predictor(cfg, 512)
    pred.load_state_dict(synth.to_torch_state(state_np), strict=True)
    return pred.to(device).eval()
