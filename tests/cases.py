"""Golden-fixture case table shared by tests/golden/make_golden.py (which runs the reference) and
the parity tests (which regenerate the same seeded inputs and compare with the stored outputs)."""
from __future__ import annotations

import os

import numpy as np

from veto_b200 import synth

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

CASES = {
    # BASELINE.json configs[0]: 1 image 592x800, 20 GT boxes -> 380 pairs, VG 151/51, PredCls
    "cfg1_predcls_vg": dict(predictor="VETOPredictor", mode="predcls", dataset="VG", n_boxes=[20],
                            batch_seed=1, weight_seed=11, spread=True, tokens=[0, 57, 379]),
    # same inputs, torch-default-like init (every pair gets the same label: vacuous for argmax but
    # pins the un-spread numerics)
    "cfg1_default_init": dict(predictor="VETOPredictor", mode="predcls", dataset="VG", n_boxes=[20],
                              batch_seed=1, weight_seed=12, spread=False),
    # ragged multi-image batch incl. the degenerate sizes: 1 box (0 pairs -> [[0,0]] placeholder), 2 boxes
    "ragged_predcls": dict(predictor="VETOPredictor", mode="predcls", dataset="VG", n_boxes=[1, 2, 7, 3],
                           batch_seed=2, weight_seed=11, spread=True, H=320, W=416, tokens=[0, 1, 2]),
    # SGDet-shaped: soft class embedding (softmax @ obj_embed), pair cap hit (12*11=132 > 100)
    "sgdet_cap": dict(predictor="VETOPredictor", mode="sgdet", dataset="VG", n_boxes=[12, 9],
                      batch_seed=3, weight_seed=13, spread=True, max_pairs=100, H=320, W=416),
    # SGDet with TEST.RELATION.REQUIRE_OVERLAP (sampling.py:38-39)
    "sgdet_overlap": dict(predictor="VETOPredictor", mode="sgdet", dataset="VG", n_boxes=[10],
                          batch_seed=4, weight_seed=13, spread=True, require_overlap=True, H=320, W=416),
    # BASELINE.json configs[3] (reduced batch): MEET group heads, GQA 201/101, divide4
    "meet_gqa": dict(predictor="VETOPredictor_MEET", mode="predcls", dataset="GQA", n_boxes=[20, 9],
                     batch_seed=5, weight_seed=14, spread=True),
    "meet_vg": dict(predictor="VETOPredictor_MEET", mode="predcls", dataset="VG", n_boxes=[10],
                    batch_seed=6, weight_seed=15, spread=True, H=320, W=416),
    # MEET at SGDet test time: obj_preds from the per-class greedy NMS (Ensemble.nms_per_cls, :3855-3874); detector
    # labels from 3 classes + jittered boxes_per_cls so that overlapping boxes compete for a class
    # MEET with EXPERT_GROUP (3 experts per group, defaults.py:864) and the voting post-processor (inference.py:93-283)
    "meet_vg_experts": dict(predictor="VETOPredictor_MEET", mode="predcls", dataset="VG", n_boxes=[10],
                            batch_seed=12, weight_seed=20, spread=True, H=320, W=416, expert_group=True, expert_noise=0.35),
    # vanilla predictor + PostProcessor at SGDet test time: late per-class NMS (obj_prediction_nms) on peaked detector
    # logits, boxes re-regressed per class (relation_head/inference.py:414-431)
    "sgdet_post_nms": dict(predictor="VETOPredictor", mode="sgdet", dataset="VG", n_boxes=[12, 7],
                           batch_seed=11, weight_seed=13, spread=True, H=320, W=416, nms_seed=4, nms_peak=3.0),
    "meet_sgdet_nms": dict(predictor="VETOPredictor_MEET", mode="sgdet", dataset="VG", n_boxes=[12, 7],
                           batch_seed=10, weight_seed=19, spread=True, H=320, W=416, nms_seed=3),
}


# Full-size fixtures at the BASELINE.json sizes: compared with the reference's outputs only (the CPU
# oracle is not re-run on them at test time: the reference itself needed minutes).  `row_stride` / `feat_stride`: the
# fixture keeps every row_stride-th logit row (argmax labels of ALL rows) and every feat_stride-th ROI channel.
FULL_CASES = {
    # BASELINE.json configs[3]: VETOPredictor_MEET PredCls, GQA 201 / 101, divide4 group heads, batch 16 x 20 boxes = 6080 pairs
    "meet_gqa_full": dict(predictor="VETOPredictor_MEET", mode="predcls", dataset="GQA", n_boxes=[20] * 16,
                          batch_seed=21, weight_seed=14, spread=True, row_stride=4, feat_stride=64),
}

# BASELINE.json configs[1]: the training step at IMS_PER_BATCH 12 x 20 GT boxes = 4560 pairs (dropout p = 0 for parity)
FULL_TRAIN_CASES = {
    "train_predcls_full": dict(predictor="VETOPredictor", mode="predcls", dataset="VG", n_boxes=[20] * 12,
                               batch_seed=22, weight_seed=16, spread=True, label_seed=74, fg_per_image=10),
}

# Training-step fixtures (tests/golden/make_golden.py run_train_case): the reference's VETOPredictor in train() mode
# with every nn.Dropout set to p = 0 (module attributes), all ordered pairs per image as gtbox_relsample yields them
# under the 1024-pair cap, seeded predicate labels; loss.backward() through the reference's own autograd.
TRAIN_CASES = {
    # BASELINE.json configs[1] shape, reduced: PredCls, ragged images, ~4 foreground pairs per image
    "train_predcls": dict(predictor="VETOPredictor", mode="predcls", dataset="VG", n_boxes=[6, 3, 5],
                          batch_seed=7, weight_seed=16, spread=True, H=320, W=416, label_seed=70, fg_per_image=4),
    # SGCls-style: soft class embedding (softmax @ obj_embed) + the constant obj_loss, class-weighted CE
    "train_sgcls": dict(predictor="VETOPredictor", mode="sgcls", dataset="VG", n_boxes=[5, 4],
                        batch_seed=8, weight_seed=17, spread=True, H=320, W=416, label_seed=71, fg_per_image=6,
                        class_weight_seed=72),
}

# MEET training fixtures (run_meet_train_case): VETOPredictor_MEET in train() mode, dropout p = 0, `random` seeded with
# sample_seed right before the forward so that the group sampling (roi_relation_predictors.py:3940-3969) is reproducible.
MEET_TRAIN_CASES = {
    "train_meet_vg": dict(predictor="VETOPredictor_MEET", mode="predcls", dataset="VG", n_boxes=[6, 4, 5],
                          batch_seed=9, weight_seed=18, spread=True, H=320, W=416, label_seed=73, fg_per_image=10,
                          sample_seed=123),
}

# RelationSampling.gtbox_relsample fixtures: seeded relation matrices; `caps` = (BATCH_SIZE_PER_IMAGE, POSITIVE_FRACTION)
RELSAMPLE_CASES = {
    "relsample_under_caps": dict(n_boxes=[20, 6, 1, 12], seed=31, fg_per_image=10, caps=(1024, 0.25)),
    "relsample_over_caps": dict(n_boxes=[20, 9, 12], seed=32, fg_per_image=14, caps=(32, 0.25)),
}

# RelationSampling.detect_relsample fixtures (synth.make_detect_case); caps = (BATCH_SIZE_PER_IMAGE, POSITIVE_FRACTION)
DETECT_SAMPLE_CASES = {
    "detsample_default": dict(seed=61, n_tgt=[9, 5, 1, 12], caps=(1024, 0.25), require_overlap=False),
    "detsample_tight": dict(seed=62, n_tgt=[10, 8], caps=(24, 0.25), require_overlap=True),
}

# recall-evaluation fixtures (SGRecall.calculate_recall of the reference on synth.make_eval_case)
EVAL_CASES = {
    "eval_recall": dict(seed=41, n_objs=[20, 9, 2, 14], n_gt_rels=12, n_pred_rels=150),
}


def case_rel_labels(c, pair_counts):
    ds = synth.VG if c["dataset"] == "VG" else synth.GQA
    return synth.make_rel_labels(c["label_seed"], pair_counts, ds["num_rel"], c["fg_per_image"])


def case_class_weight(c):
    if "class_weight_seed" not in c:
        return None
    ds = synth.VG if c["dataset"] == "VG" else synth.GQA
    return np.random.default_rng(c["class_weight_seed"]).uniform(0.5, 2.0, ds["num_rel"]).astype(np.float32)


def case_batch(c, features=True):
    ds = synth.VG if c["dataset"] == "VG" else synth.GQA
    batch = synth.make_batch(c["batch_seed"], c["n_boxes"], H=c.get("H", 592), W=c.get("W", 800),
                             num_obj=ds["num_obj"], mode=c["mode"], features=features)
    if "nms_seed" in c:
        synth.add_nms_fields(batch, c["nms_seed"], peak=c.get("nms_peak", 0.0))
    return batch


def case_state(c):
    ds = synth.VG if c["dataset"] == "VG" else synth.GQA
    if c["predictor"].endswith("MEET"):
        return synth.meet_state(c["weight_seed"], ds["num_obj"], synth.GROUP_SPLITS[(c["dataset"], "divide4")],
                                spread=c["spread"], experts_per_group=3 if c.get("expert_group") else 1,
                                expert_group=bool(c.get("expert_group")), expert_noise=c.get("expert_noise", 0.0))
    return synth.predictor_state(c["weight_seed"], ds["num_obj"], ds["num_rel"], spread=c["spread"])


def load_golden(name):
    return np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
