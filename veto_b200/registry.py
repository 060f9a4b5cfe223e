"""Plugin registries with the reference's names (pysgg/modeling/registry.py, pysgg/utils/registry.py:4-45).

``ROI_RELATION_PREDICTOR["VETOPredictor" | "VETOPredictor_MEET"]`` and
``ROI_BOX_FEATURE_EXTRACTORS["VETOFeatureExtractor"]`` and ``BACKBONES["R-18-C4"]`` (the depth backbone) resolve to the
B200 drop-ins.  With the reference
importable, ``install_into_reference()`` replaces the entries of pysgg's own registries, so that
``cfg.MODEL.ROI_RELATION_HEAD.PREDICTOR`` selects them unchanged
(roi_relation_predictors.py:4152-4154; roi_box_feature_extractors.py:315-323).
"""
from __future__ import annotations


class Registry(dict):
    """dict with a decorator-style ``register`` (pysgg/utils/registry.py:34-45)."""

    def register(self, module_name, module=None):
        if module is not None:
            assert module_name not in self, f"{module_name} already registered"
            self[module_name] = module
            return module

        def register_fn(fn):
            assert module_name not in self, f"{module_name} already registered"
            self[module_name] = fn
            return fn

        return register_fn


ROI_RELATION_PREDICTOR = Registry()
ROI_BOX_FEATURE_EXTRACTORS = Registry()
BACKBONES = Registry()          # "R-18-C4": the depth backbone (backbone/backbone.py:83-93)


def install_into_reference() -> bool:
    """Point the reference's registries at the drop-ins.  Returns False when pysgg is not importable.

    The reference's Registry.register asserts the name is unused (utils/registry.py:4-6), so the
    entries are replaced by item assignment, as SURVEY.md §8b prescribes."""
    try:
        from pysgg.modeling import registry as ref_registry
    except Exception:
        return False
    from . import depth_backbone, feature_extractor, predictor  # noqa: F401  (registers into the local registries)
    for name, cls in ROI_RELATION_PREDICTOR.items():
        ref_registry.ROI_RELATION_PREDICTOR[name] = cls
    for name, cls in ROI_BOX_FEATURE_EXTRACTORS.items():
        ref_registry.ROI_BOX_FEATURE_EXTRACTORS[name] = cls
    for name, fn in BACKBONES.items():
        ref_registry.BACKBONES[name] = fn
    return True


def make_roi_relation_predictor(cfg, in_channels):
    """roi_relation_predictors.py:4152-4154."""
    from . import predictor  # noqa: F401
    return ROI_RELATION_PREDICTOR[cfg.MODEL.ROI_RELATION_HEAD.PREDICTOR](cfg, in_channels)


def make_roi_box_feature_extractor(cfg, in_channels, half_out=False, cat_all_levels=False, for_relation=False):
    """roi_box_feature_extractors.py:315-323 (the relation-head branch with DATASETS.USE_DEPTH)."""
    from . import feature_extractor  # noqa: F401
    name = cfg.MODEL.ROI_RELATION_HEAD.FEATURE_EXTRACTOR_MINI
    return ROI_BOX_FEATURE_EXTRACTORS[name](cfg, in_channels, half_out, cat_all_levels, for_relation)
