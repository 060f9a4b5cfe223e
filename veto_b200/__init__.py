"""veto_b200 — B200-native (sm_100a) implementation of the VETO relation-prediction hot path of visinf/veto.

Public surface (mirrors the reference's plugin API for this path):

* ``registry.ROI_RELATION_PREDICTOR["VETOPredictor" | "VETOPredictor_MEET"]``, ``registry.ROI_BOX_FEATURE_EXTRACTORS
  ["VETOFeatureExtractor"]``, ``registry.install_into_reference()``;
* ``sampling.RelationSampling.prepare_test_pairs``, ``postprocess.PostProcessor``;
* ``ops`` — functional wrappers over the C ABI of ``libveto_b200.so`` (``include/veto_b200.h``).

Importing the package does not need a GPU; running any op does (there is no CPU fallback).
"""
from . import config, lib, registry, structures  # noqa: F401

__version__ = "0.1.0"


def load_modules():
    """Import the drop-in modules (registers them) and return the registries."""
    from . import depth_backbone, feature_extractor, postprocess, predictor, relation_head, sampling  # noqa: F401
    return registry.ROI_RELATION_PREDICTOR, registry.ROI_BOX_FEATURE_EXTRACTORS
