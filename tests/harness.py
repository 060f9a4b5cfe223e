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


from veto_b200.workloads import (boxlists, build_predictor, device_features, make_cfg,  # noqa: F401,E402
                                 onehot_logits)


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
