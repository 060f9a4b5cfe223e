"""The configuration keys the VETO relation-head path reads, with the values of the reference's
``pysgg/config/defaults.py`` overridden by ``configs/VETO_final.yaml`` (SURVEY.md §5 "Config / flags").

The drop-in modules accept the reference's own yacs ``CfgNode`` (attribute access is all they use);
``default_cfg()`` builds an equivalent attribute dict for use without the reference.
"""
from __future__ import annotations

import copy


class CfgNode(dict):
    """Attribute-access dict with the subset of the yacs CfgNode API the path needs."""

    def __init__(self, init=None):
        super().__init__()
        for k, v in (init or {}).items():
            self[k] = CfgNode(v) if isinstance(v, dict) and not isinstance(v, CfgNode) else v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v

    def clone(self):
        return copy.deepcopy(self)

    def merge_from_list(self, lst):
        for key, v in zip(lst[0::2], lst[1::2]):
            node = self
            parts = key.split(".")
            for p in parts[:-1]:
                node = node[p]
            if parts[-1] not in node:
                raise KeyError(f"unknown config key {key}")
            node[parts[-1]] = v
        return self

    def freeze(self):
        pass

    def defrost(self):
        pass


_DEFAULTS = {
    "MODEL": {
        "DEVICE": "cuda",
        "DEPTH_BACKBONE": {"CONV_BODY": "R-18-C4"},             # defaults.py:117-123
        "ROI_BOX_HEAD": {
            "POOLER_RESOLUTION": 7,
            "POOLER_SCALES": (0.25, 0.125, 0.0625, 0.03125),   # VETO_final.yaml:39
            "POOLER_SAMPLING_RATIO": 2,
            "VG_NUM_CLASSES": 151,                              # defaults.py:240
            "GQA_200_NUM_CLASSES": 201,                         # defaults.py:241
        },
        "ROI_RELATION_HEAD": {
            "PREDICTOR": "VETOPredictor",
            "FEATURE_EXTRACTOR_MINI": "VETOFeatureExtractor",
            "USE_GT_BOX": True,
            "USE_GT_OBJECT_LABEL": True,
            "POOLER_RESOLUTION": 8,                             # VETO_final.yaml:57
            "MAX_PROPOSAL_PAIR": 2048,                          # defaults.py:305
            "BATCH_SIZE_PER_IMAGE": 1024,                       # VETO_final.yaml:64
            "POSITIVE_FRACTION": 0.25,                          # VETO_final.yaml:65
            "CONTEXT_HIDDEN_DIM": 512,
            "CONTEXT_POOLING_DIM": 4096,
            "VG_NUM_CLASSES": 51,                               # defaults.py:301
            "GQA_200_NUM_CLASSES": 101,                         # defaults.py:302
            "VETOTRANSFORMER": {"PATCH_SIZE": 2, "T_INPUT_DIM": 576, "ENC_LAYERS": 6, "NHEADS": 6,
                                "EMB_DROPOUT": 0.35, "T_DROPOUT": 0.35},
        },
    },
    "TEST": {"RELATION": {"REQUIRE_OVERLAP": False, "LATER_NMS_PREDICTION_THRES": 0.5}},
    "DATASETS": {"USE_DEPTH": True},
    "GLOBAL_SETTING": {"DATASET_CHOICE": "VG", "USE_BIAS": False, "BETA_LOSS": False},
    "GCL_SETTING": {"GROUP_SPLIT_MODE": "divide4", "ZERO_LABEL_PADDING_MODE": "rand_insert"},
    "ENSEMBLE_LEARNING": {"ENABLED": False, "TYPE": "group", "VOTING": "C", "EXPERT_GROUP": False},
    "GLOVE_DIR": "",
    "OUTPUT_DIR": "",
    # extension (not a reference key): arithmetic of the encoder GEMMs, see include/veto_b200.h
    "VETO_B200": {"PRECISION": "f16c8", "CHUNK_PAIRS": 0, "FREQ_BIAS": False, "PRED_COUNTS": "",
                  "MEET_REFERENCE_DRAWS": False},
}

# predicate_stage_count of SHA_GCL_extra/group_chosen_function.py:6-95 (groups are contiguous id ranges)
GROUP_SPLITS = {
    ("VG", "divide3"): [3, 3, 8, 6, 20, 10],
    ("VG", "divide4"): [4, 6, 9, 19, 12],
    ("VG", "divide5"): [4, 8, 10, 28],
    ("VG", "divide7new"): [2, 4, 5, 6, 8, 10, 15],
    ("VG", "average"): [10, 10, 10, 10, 10],
    ("GQA", "divide3"): [4, 4, 11, 16, 31, 34],
    ("GQA", "divide4"): [5, 10, 20, 65],
    ("GQA", "divide5"): [7, 14, 28, 51],
    ("GQA", "average"): [20, 20, 20, 20, 20],
}


def default_cfg() -> CfgNode:
    return CfgNode(copy.deepcopy(_DEFAULTS))


def get(cfg, path: str, default=None):
    """cfg.A.B.C with a default when a node is missing (the extension keys are absent from a reference cfg)."""
    node = cfg
    for p in path.split("."):
        try:
            node = getattr(node, p)
        except (AttributeError, KeyError):
            return default
    return node


def num_classes(cfg):
    """(num_obj, num_rel) for cfg.GLOBAL_SETTING.DATASET_CHOICE."""
    ds = get(cfg, "GLOBAL_SETTING.DATASET_CHOICE", "VG")
    if ds == "GQA":
        return cfg.MODEL.ROI_BOX_HEAD.GQA_200_NUM_CLASSES, cfg.MODEL.ROI_RELATION_HEAD.GQA_200_NUM_CLASSES
    return cfg.MODEL.ROI_BOX_HEAD.VG_NUM_CLASSES, cfg.MODEL.ROI_RELATION_HEAD.VG_NUM_CLASSES
