"""Shared helpers of the GPU parity tests: move a synthetic batch (veto_b200.synth) to the device, build the
BoxLists the drop-in modules take, and run the whole head through the public (reference-shaped) API."""
from __future__ import annotations

import numpy as np
import torch

from veto_b200 import config as vcfg
from veto_b200 import registry, synth
from veto_b200.postprocess import make_roi_relation_post_processor
from veto_b200.sampling import make_roi_relation_samp_processor
from veto_b200.structures import BoxList


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


def build_predictor(cfg, state_np, device):
    pred = registry.make_roi_relation_predictor(cfg, 512)
    pred.load_state_dict(synth.to_torch_state(state_np), strict=True)
    return pred.to(device).eval()


def run_head(cfg, state_np, batch, device, post=True):
    """prepare_test_pairs -> VETOFeatureExtractor -> predictor (-> PostProcessor), as ROIRelationHead.forward
    (relation_head.py:134-243) strings them together."""
    num_obj = vcfg.num_classes(cfg)[0]
    bls = boxlists(batch, device, num_obj)
    feats, depth = device_features(batch, device)
    samp = make_roi_relation_samp_processor(cfg)
    fe = registry.make_roi_box_feature_extractor(cfg, 256, for_relation=True).to(device).eval()
    pred = build_predictor(cfg, state_np, device)
    with torch.no_grad():
        pairs = samp.prepare_test_pairs(device, bls)
        x2d, d2d, _, _ = fe(feats, bls, depth_features=depth)
        obj_d, rel_d, losses, incre, chosen, custom = pred(bls, pairs, None, None, roi_features=x2d,
                                                           roi_depth_features=d2d)
        res = None
        if post and not isinstance(rel_d, dict) and batch["mode"] == "predcls":
            pp = make_roi_relation_post_processor(cfg)
            res = pp((rel_d, [b.get_field("predict_logits") for b in bls]), pairs, bls)
    return dict(pairs=pairs, x2d=x2d, d2d=d2d, obj_dists=obj_d, rel_dists=rel_d, incre=incre, results=res,
                predictor=pred, boxlists=bls)


def np_(t):
    return t.detach().cpu().numpy()
