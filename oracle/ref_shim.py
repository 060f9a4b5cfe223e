"""ORACLE tooling — import the UNMODIFIED reference (/root/reference) on CPU under import shims.

Works only in the build container (the reference tree does not exist on the GPU box).  Used by
tests/golden/make_golden.py to produce the committed golden fixtures, and by tests marked
`needs_reference` (skipped when /root/reference is absent).  Recipe: SURVEY.md Appendix A.
Nothing in the reference tree is edited or copied; only module attributes are patched at run time.
"""
from __future__ import annotations

import ast
import copy
import os
import sys
import types

REF_ROOT = os.environ.get("VETO_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "pysgg"))


class _Stub(types.ModuleType):
    """Inert module/object: any attribute is another stub; calling it with one callable returns the
    callable (so decorators become identity); usable as a base class."""

    def __init__(self, name="stub"):
        super().__init__(name)
        self.__path__ = []

    def __getattr__(self, k):
        if k.startswith("__") and k.endswith("__"):
            raise AttributeError(k)
        s = _Stub(f"{self.__name__}.{k}")
        setattr(self, k, s)
        return s

    def __call__(self, *a, **k):
        if len(a) == 1 and callable(a[0]) and not k:
            return a[0]
        return _Stub(self.__name__ + "()")

    def __mro_entries__(self, bases):
        return (object,)


class CfgNode(dict):
    """Functional stand-in for yacs.config.CfgNode (attribute dict + merge helpers)."""

    def __init__(self, init=None, new_allowed=False, **kw):
        super().__init__()
        if init:
            for k, v in init.items():
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

    def freeze(self):
        pass

    def defrost(self):
        pass

    @staticmethod
    def _coerce(v, old=None):
        if isinstance(v, str):
            try:
                lit = ast.literal_eval(v)
                if isinstance(lit, (tuple, list)) or (old is not None and not isinstance(old, str)):
                    v = lit
            except (ValueError, SyntaxError):
                pass
        if isinstance(old, tuple) and isinstance(v, list):
            v = tuple(v)
        if isinstance(old, float) and isinstance(v, int):
            v = float(v)
        return v

    def _merge(self, other):
        for k, v in other.items():
            if isinstance(v, dict):
                if k not in self or not isinstance(self[k], CfgNode):
                    self[k] = CfgNode()
                self[k]._merge(v)
            else:
                self[k] = self._coerce(v, self.get(k))

    def merge_from_file(self, path):
        import yaml
        with open(path) as f:
            self._merge(yaml.safe_load(f))

    def merge_from_list(self, lst):
        for key, v in zip(lst[0::2], lst[1::2]):
            node = self
            parts = key.split(".")
            for p in parts[:-1]:
                node = node[p]
            node[parts[-1]] = self._coerce(v, node.get(parts[-1]))


_loaded = None


def load():
    """Import the reference modules; returns a namespace of the handles the harness needs."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    import torch
    import torchvision

    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    torch._six = types.SimpleNamespace(PY37=True, PY3=True, string_classes=(str,), int_classes=(int,))
    sys.modules["torch._six"] = torch._six
    import torchvision.models.resnet as _tvr
    if not hasattr(_tvr, "model_urls"):
        _tvr.model_urls = {}
    for name in ["ipdb", "matplotlib", "matplotlib.pyplot", "matplotlib.image", "pysgg._C", "apex",
                 "pycocotools", "pycocotools.mask", "pycocotools.coco", "pycocotools.cocoeval", "h5py", "cv2", "tensorboardX", "termcolor",
                 "overrides", "graphviz", "gpustat"]:
        try:
            __import__(name)
        except Exception:
            sys.modules[name] = _Stub(name)
    yacs = types.ModuleType("yacs")
    yacs_config = types.ModuleType("yacs.config")
    yacs_config.CfgNode = CfgNode
    yacs.config = yacs_config
    sys.modules.setdefault("yacs", yacs)
    sys.modules.setdefault("yacs.config", yacs_config)

    import pysgg.modeling.roi_heads.relation_head.roi_relation_predictors as P
    from pysgg.config import cfg
    from pysgg.structures.bounding_box import BoxList
    from pysgg.modeling.roi_heads.box_head.roi_box_feature_extractors import make_roi_box_feature_extractor
    from pysgg.modeling.roi_heads.relation_head.sampling import make_roi_relation_samp_processor
    from pysgg.modeling.roi_heads.relation_head.inference import make_roi_relation_post_processor
    from pysgg.modeling.roi_heads.relation_head.model_kern import to_onehot

    class _C:  # replaces pysgg._C inside layers/roi_align.py with the bit-identical torchvision op
        @staticmethod
        def roi_align_forward(x, rois, s, ph, pw, sr):
            return torchvision.ops.roi_align(x, rois, (ph, pw), s, sr, aligned=False)

        @staticmethod
        def roi_align_backward(g, rois, s, ph, pw, bs, ch, h, w, sr):
            return torch.ops.torchvision._roi_align_backward(g, rois, s, ph, pw, bs, ch, h, w, sr, False)

    sys.modules["pysgg.layers.roi_align"]._C = _C

    def names(n, kind):
        return ["__background__"] + [f"{kind}{i}" for i in range(1, n)]

    ns = types.SimpleNamespace(P=P, cfg=cfg, BoxList=BoxList, torch=torch, to_onehot=to_onehot,
                               make_extractor=make_roi_box_feature_extractor,
                               make_sampler=make_roi_relation_samp_processor,
                               make_post=make_roi_relation_post_processor, names=names)
    _loaded = ns
    return ns


def make_cfg(ns, predictor="VETOPredictor", mode="predcls", dataset="VG", max_pairs=2048,
             require_overlap=False, expert_group=False):
    cfg = ns.cfg.clone()
    cfg.merge_from_file(os.path.join(REF_ROOT, "configs/VETO_final.yaml"))
    cfg.merge_from_list([
        "MODEL.ROI_RELATION_HEAD.PREDICTOR", predictor,
        "MODEL.ROI_RELATION_HEAD.USE_GT_BOX", mode in ("predcls", "sgcls"),
        "MODEL.ROI_RELATION_HEAD.USE_GT_OBJECT_LABEL", mode == "predcls",
        "MODEL.ROI_RELATION_HEAD.MAX_PROPOSAL_PAIR", max_pairs,
        "TEST.RELATION.REQUIRE_OVERLAP", require_overlap,
        "GLOBAL_SETTING.BETA_LOSS", False,
        "GLOBAL_SETTING.DATASET_CHOICE", dataset,
        "ENSEMBLE_LEARNING.ENABLED", predictor.endswith("MEET"),
        "ENSEMBLE_LEARNING.EXPERT_GROUP", expert_group,
        "MODEL.DEVICE", "cpu",
    ])
    return cfg


def build_predictor(ns, cfg, num_obj, num_rel, state=None):
    """Construct the reference predictor for `cfg` and load a synthetic state_dict."""
    P, torch = ns.P, ns.torch
    P.get_dataset_statistics = lambda c: {"obj_classes": ns.names(num_obj, "obj"),
                                          "rel_classes": ns.names(num_rel, "rel")}
    P.obj_edge_vectors = lambda names, wv_dir, wv_dim: torch.randn(len(names), wv_dim)
    P.cfg = cfg                      # VETOPredictor_MEET reads the *global* cfg (:3902-3904)
    pred = P.registry.ROI_RELATION_PREDICTOR[cfg.MODEL.ROI_RELATION_HEAD.PREDICTOR](cfg, 512)
    if state is not None:
        missing, unexpected = pred.load_state_dict(state, strict=True), None
    return pred.eval()


def make_boxlists(ns, batch, num_obj):
    """BoxLists with the fields ROIRelationHead.forward sets (relation_head.py:104-111)."""
    torch = ns.torch
    out = []
    for i in range(batch["B"]):
        bl = ns.BoxList(torch.from_numpy(batch["boxes"][i]), (batch["W"], batch["H"]), mode="xyxy")
        bl.add_field("labels", torch.from_numpy(batch["labels"][i]))
        if batch["mode"] == "predcls":
            lab = torch.from_numpy(batch["labels"][i])
            bl.add_field("predict_logits", ns.to_onehot(lab, num_obj))
            bl.add_field("pred_scores", torch.ones(len(lab)))
            bl.add_field("pred_labels", lab)
        else:
            bl.add_field("predict_logits", torch.from_numpy(batch["predict_logits"][i]))
            bl.add_field("pred_scores", torch.from_numpy(batch["pred_scores"][i]))
            bl.add_field("pred_labels", torch.from_numpy(batch["pred_labels"][i]))
            if "boxes_per_cls" in batch:
                bpc = torch.from_numpy(batch["boxes_per_cls"][i])
            else:
                bpc = torch.from_numpy(batch["boxes"][i])[:, None, :].expand(-1, num_obj, -1).contiguous()
            bl.add_field("boxes_per_cls", bpc)
        out.append(bl)
    return out
