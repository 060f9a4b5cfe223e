"""The drop-in claim guarded in-tree: with the reference importable, ``install_into_reference()``
makes the reference's OWN ``build_roi_relation_head(cfg, 256)`` (relation_head.py:251-257, :52-64) construct the B200
classes, and a state_dict taken from a reference-constructed predictor loads strictly into them.  CPU only (construction
and state-dict plumbing: no kernel runs); needs /root/reference, i.e. the build container."""
import pytest
import torch

pytestmark = pytest.mark.needs_reference


@pytest.fixture()
def ref_ns():
    from oracle import ref_shim
    ns = ref_shim.load()
    from pysgg.modeling import registry as ref_registry
    saved = {r: dict(getattr(ref_registry, r)) for r in ("ROI_RELATION_PREDICTOR", "ROI_BOX_FEATURE_EXTRACTORS", "BACKBONES")}
    yield ns
    for r, entries in saved.items():        # other tests in this process still want the reference's own classes
        reg = getattr(ref_registry, r)
        reg.clear()
        reg.update(entries)


@pytest.mark.parametrize("predictor,dataset,num_obj,num_rel", [("VETOPredictor", "VG", 151, 51),
                                                               ("VETOPredictor_MEET", "GQA", 201, 101)])
def test_reference_head_builds_the_dropins(ref_ns, predictor, dataset, num_obj, num_rel):
    from oracle import ref_shim
    from veto_b200 import feature_extractor as vfe
    from veto_b200 import predictor as vpred
    from veto_b200 import registry as vreg
    ns = ref_ns
    cfg = ref_shim.make_cfg(ns, predictor=predictor, dataset=dataset)
    torch.manual_seed(0)
    reference_pred = ref_shim.build_predictor(ns, cfg, num_obj, num_rel)          # the reference's own class
    assert type(reference_pred).__module__.startswith("pysgg.")
    ref_state = reference_pred.state_dict()

    assert vreg.install_into_reference() is True
    # the hooks the reference resolves from dataset files (absent here), for the drop-in as for the reference
    stats = {"obj_classes": ns.names(num_obj, "obj"), "rel_classes": ns.names(num_rel, "rel")}
    vpred.get_dataset_statistics = lambda c: stats
    vpred.obj_edge_vectors = lambda names, wv_dir, wv_dim: torch.randn(len(names), wv_dim)
    from pysgg.modeling.roi_heads.relation_head.relation_head import build_roi_relation_head
    head = build_roi_relation_head(cfg, 256)
    assert type(head).__module__.startswith("pysgg.")                             # the reference's ROIRelationHead ...
    assert isinstance(head.predictor, vpred.ROI_RELATION_PREDICTOR[predictor])    # ... holding the B200 predictor
    assert type(head.predictor).__module__ == "veto_b200.predictor"
    assert isinstance(head.box_feature_extractor, vfe.VETOFeatureExtractor)
    assert head.box_feature_extractor.out_channels == 256
    # same parameter / buffer names, shapes and dtypes: a reference checkpoint loads strictly
    mine = head.predictor.state_dict()
    assert list(mine) == list(ref_state)
    for k, v in ref_state.items():
        assert mine[k].shape == v.shape and mine[k].dtype == v.dtype, k
    head.predictor.load_state_dict(ref_state, strict=True)
    for k, v in head.predictor.state_dict().items():
        assert torch.equal(v, ref_state[k]), k
    # the depth backbone entry of the reference's BACKBONES registry
    from pysgg.modeling import registry as ref_registry
    model = ref_registry.BACKBONES["R-18-C4"](cfg, True)
    assert type(model.body).__module__ == "veto_b200.depth_backbone" and model.out_channels == 256
    # without CUDA the drop-in refuses to compute instead of falling back
    bls = ref_shim.make_boxlists(ns, {"B": 1, "W": 64, "H": 64, "mode": "predcls",
                                     "boxes": [torch.tensor([[1., 1., 20., 20.], [5., 5., 40., 40.]]).numpy()],
                                     "labels": [torch.tensor([1, 2]).numpy()]}, num_obj)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            head.predictor.eval()(bls, [torch.tensor([[0, 1], [1, 0]])], None, None,
                                  roi_features=torch.zeros(2, 256, 8, 8), roi_depth_features=torch.zeros(2, 256, 8, 8))
