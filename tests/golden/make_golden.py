"""Generate the golden fixtures by running the UNMODIFIED reference (/root/reference) on CPU.

Run in the build container only:  python tests/golden/make_golden.py
Inputs and weights are regenerated from seeds by veto_b200.synth at test time (they are too large
to commit: 43 MB of feature maps per image, 70 MB of weights); the fixtures hold the reference's
OUTPUTS plus sha256 digests of the inputs/weights they were computed from.
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from veto_b200 import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

# name -> case description; shared with tests/cases.py
from tests.cases import CASES, case_batch, case_state  # noqa: E402


def digest(arrs) -> str:
    h = hashlib.sha256()
    for a in arrs:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def run_case(ns, name, c):
    import torch
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    ds = synth.VG if c["dataset"] == "VG" else synth.GQA
    meet = c["predictor"].endswith("MEET")
    cfg = ref_shim.make_cfg(ns, predictor=c["predictor"], mode=c["mode"], dataset=c["dataset"],
                            max_pairs=c.get("max_pairs", 2048), require_overlap=c.get("require_overlap", False),
                            expert_group=bool(c.get("expert_group")))
    sd = case_state(c)
    pred = ref_shim.build_predictor(ns, cfg, ds["num_obj"], ds["num_rel"], synth.to_torch_state(sd))
    batch = case_batch(c)
    bls = ref_shim.make_boxlists(ns, batch, ds["num_obj"])
    fe = ns.make_extractor(cfg, 256, for_relation=True).eval()
    samp, post = ns.make_sampler(cfg), ns.make_post(cfg)
    out = {}
    if "nms_seed" in c and meet:
        # record what Ensemble.nms_per_cls saw and returned (its softmax input is computed per image on CPU torch)
        real_nms = pred.model.nms_per_cls
        def _nms(obj_dists, boxes_per_cls, num_objs):
            out["nms_scores"] = np.concatenate([torch.softmax(d, -1).numpy() for d in obj_dists.split(num_objs, dim=0)])
            res = real_nms(obj_dists, boxes_per_cls, num_objs)
            out["nms_labels"] = res.numpy().copy()
            return res
        pred.model.nms_per_cls = _nms
    with torch.no_grad():
        pairs = samp.prepare_test_pairs(torch.device("cpu"), bls)
        feats = [torch.from_numpy(f) for f in batch["feats"]]
        feats.append(torch.zeros(batch["B"], 256, 1, 1))            # P6: present in the reference, unused
        x2d, d2d, _, _ = fe(feats, bls, depth_features=torch.from_numpy(batch["depth"]))
        obj_d, rel_d, losses, incre, chosen, custom = pred(bls, pairs, None, None, roi_features=x2d,
                                                          roi_depth_features=d2d)
    if "nms_seed" in c and meet:
        changed = int((out["nms_labels"] != np.concatenate(batch["pred_labels"])).sum())
        print(f"{name}: per-class NMS changed {changed} of {len(out['nms_labels'])} labels")
        assert changed >= 2, "the NMS case must exercise suppression"
    out["input_digest"] = np.array(digest(batch["feats"] + [batch["depth"]] + batch["boxes"] + batch["labels"]))
    out["weight_digest"] = np.array(digest([sd[k] for k in sorted(sd)]))
    out["pairs"] = np.concatenate([p.numpy() for p in pairs])
    out["pair_counts"] = np.array([len(p) for p in pairs])
    fs = c.get("feat_stride", 16)
    out["x2d_sub"] = x2d.numpy()[:, ::fs].copy()
    out["d2d_sub"] = d2d.numpy()[:, ::fs].copy()
    out["x2d_sum"] = x2d.numpy().sum(axis=(1, 2, 3), dtype=np.float64)
    out["d2d_sum"] = d2d.numpy().sum(axis=(1, 2, 3), dtype=np.float64)
    out["obj_dists_argmax"] = np.concatenate([o.numpy().argmax(1) for o in obj_d])
    if meet:
        rs = c.get("row_stride", 1)
        for k, v in rel_d.items():
            out["logits_" + k] = v.numpy()[::rs].copy()
            if rs > 1:      # full-size case: every rs-th logit row, the argmax label (classes 1..) of every row
                out["argmax_" + k] = v.numpy()[:, 1:].argmax(1).astype(np.int16)
                top2 = np.sort(v.numpy()[:, 1:], 1)[:, -2:]
                out["margin_" + k] = (top2[:, 1] - top2[:, 0]).astype(np.float32)
        out["incre_idx_list"] = np.array(incre)
        if c["mode"] == "predcls" and batch["B"] == 1 and c.get("expert_group"):
            # EXPERT_GROUP voting branch (inference.py:93-283), both voting rules on the same logits
            real_cuda = torch.Tensor.cuda
            torch.Tensor.cuda = lambda self, *a, **k: self
            try:
                for voting in ("C", "U"):
                    cfg.merge_from_list(["ENSEMBLE_LEARNING.VOTING", voting])
                    vpost = ns.make_post(cfg)
                    import copy
                    vb = [copy.deepcopy(b) for b in bls]
                    res = vpost((rel_d, [b.get_field("predict_logits") for b in vb]), pairs, vb, incre_idx_list=incre)
                    r0 = res[0]
                    out[f"vote{voting}_pairs"] = r0.get_field("rel_pair_idxs").numpy()
                    out[f"vote{voting}_probs"] = r0.get_field("pred_rel_scores").numpy()
                    out[f"vote{voting}_labels"] = r0.get_field("pred_rel_labels").numpy()
                    print(f"{name}: voting {voting}: {len(out[f'vote{voting}_labels'])} of {5 * len(pairs[0])} candidates survive")
            finally:
                torch.Tensor.cuda = real_cuda
        elif c["mode"] == "predcls" and batch["B"] == 1:
            # MEET 'ensemble' branch of the post-processor (inference.py:284-397); its hard-coded .cuda() calls are
            # neutralised for the CPU run, nothing else is touched
            real_cuda = torch.Tensor.cuda
            torch.Tensor.cuda = lambda self, *a, **k: self
            try:
                res = post((rel_d, [b.get_field("predict_logits") for b in bls]), pairs, bls, incre_idx_list=incre)
            finally:
                torch.Tensor.cuda = real_cuda
            r0 = res[0]
            out["mpost_pairs"] = r0.get_field("rel_pair_idxs").numpy()
            out["mpost_probs"] = r0.get_field("pred_rel_scores").numpy()
            out["mpost_labels"] = r0.get_field("pred_rel_labels").numpy()
    else:
        out["logits"] = np.concatenate([r.numpy() for r in rel_d])
        if c["mode"] == "sgdet" and "nms_seed" in c:
            # SGDet branch of the vanilla post-processor (inference.py:414-431): late NMS + per-class box regression
            cfg.merge_from_list(["TEST.RELATION.LATER_NMS_PREDICTION_THRES", 0.5])
            post = ns.make_post(cfg)
            obj_logits = [b.get_field("predict_logits") for b in bls]
            out["post_obj_prob"] = np.concatenate([torch.softmax(l, -1).numpy() for l in obj_logits])
            res = post((rel_d, obj_logits), pairs, bls)
            out["post_obj_labels"] = np.concatenate([r.get_field("pred_labels").numpy() for r in res])
            out["post_obj_scores"] = np.concatenate([r.get_field("pred_scores").numpy() for r in res])
            out["post_boxes"] = np.concatenate([r.bbox.numpy() for r in res])
            out["post_pairs"] = np.concatenate([r.get_field("rel_pair_idxs").numpy() for r in res])
            out["post_labels"] = np.concatenate([r.get_field("pred_rel_labels").numpy() for r in res])
            out["post_scores"] = np.concatenate([r.get_field("pred_rel_scores").numpy()[:, 1:].max(1) for r in res])
            changed = int((out["post_obj_labels"] != np.concatenate(batch["pred_labels"])).sum())
            print(f"{name}: late NMS changed {changed} of {len(out['post_obj_labels'])} labels")
            assert changed >= 2
        if c["mode"] == "predcls":
            res = post((rel_d, [b.get_field("predict_logits") for b in bls]), pairs, bls)
            out["post_pairs"] = np.concatenate([r.get_field("rel_pair_idxs").numpy() for r in res])
            out["post_labels"] = np.concatenate([r.get_field("pred_rel_labels").numpy() for r in res])
            out["post_scores"] = np.concatenate([r.get_field("pred_rel_scores").numpy()[:, 1:].max(1) for r in res])
    if c.get("tokens"):
        # token tensor of a few pairs, captured at the encoder input (model_veto.py:16)
        keep = {}
        tr = pred.fusion_transformer.transformer if not meet else pred.model.fusion_transformer.transformer
        def _hook(m, i, o):
            keep["tok"] = o.detach().numpy()
        h = tr.register_forward_hook(_hook)
        with torch.no_grad():
            pred(bls, pairs, None, None, roi_features=x2d, roi_depth_features=d2d)
        h.remove()
        sel = np.array(c["tokens"])
        out["token_rows"] = sel
        out["tokens"] = keep["tok"][sel]
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: wrote {path} ({os.path.getsize(path) / 1024:.0f} KB)",
          {k: v.shape for k, v in out.items() if hasattr(v, 'shape') and v.ndim})


def grad_summary(g: np.ndarray):
    """What a fixture keeps of one gradient tensor: its norm and sum in float64 plus 64 fixed samples."""
    flat = g.reshape(-1).astype(np.float64)
    idx = np.unique(np.concatenate([np.arange(min(32, flat.size)), np.linspace(0, flat.size - 1, 32).astype(np.int64)]))
    return np.array([np.sqrt((flat ** 2).sum()), flat.sum()]), idx, flat[idx].astype(np.float32)


def run_train_case(ns, name, c):
    """One training step of the UNMODIFIED reference predictor (train() mode, dropout p = 0): rel_loss and its
    gradients wrt every parameter and wrt the depth feature map (through the reference's _ROIAlign autograd)."""
    import torch
    from tests.cases import case_class_weight, case_rel_labels
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    ds = synth.VG if c["dataset"] == "VG" else synth.GQA
    cfg = ref_shim.make_cfg(ns, predictor=c["predictor"], mode=c["mode"], dataset=c["dataset"])
    sd = case_state(c)
    cw = case_class_weight(c)
    if cw is not None:
        sd = dict(sd)
        sd["criterion_loss_rel.weight"] = cw
    pred = ref_shim.build_predictor(ns, cfg, ds["num_obj"], ds["num_rel"], synth.to_torch_state(sd)).train()
    for m in pred.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    batch = case_batch(c)
    bls = ref_shim.make_boxlists(ns, batch, ds["num_obj"])
    fe = ns.make_extractor(cfg, 256, for_relation=True).train()
    samp = ns.make_sampler(cfg)
    pairs = samp.prepare_test_pairs(torch.device("cpu"), bls)
    rel_labels = [torch.from_numpy(l) for l in case_rel_labels(c, [len(p) for p in pairs])]
    feats = [torch.from_numpy(f) for f in batch["feats"]]
    feats.append(torch.zeros(batch["B"], 256, 1, 1))
    depth = torch.from_numpy(batch["depth"]).clone().requires_grad_(True)
    x2d, d2d, _, _ = fe(feats, bls, depth_features=depth)
    d2d.retain_grad()
    obj_d, rel_d, losses, incre, chosen, custom = pred(bls, pairs, rel_labels, None, roi_features=x2d,
                                                       roi_depth_features=d2d)
    assert obj_d is None and rel_d is None
    losses["rel_loss"].backward()
    out = {"input_digest": np.array(digest(batch["feats"] + [batch["depth"]] + batch["boxes"] + batch["labels"])),
           "weight_digest": np.array(digest([sd[k] for k in sorted(sd)])),
           "rel_labels": np.concatenate([l.numpy() for l in rel_labels]),
           "pair_counts": np.array([len(p) for p in pairs]),
           "rel_loss": np.array(losses["rel_loss"].item(), np.float64)}
    if "obj_loss" in losses:
        out["obj_loss"] = np.array(losses["obj_loss"].item(), np.float64)
    no_grad = []
    for k, p in pred.named_parameters():
        if p.grad is None:
            no_grad.append(k)
            continue
        stat, idx, val = grad_summary(p.grad.numpy())
        out["gstat/" + k], out["gidx/" + k], out["gval/" + k] = stat, idx, val
    out["no_grad"] = np.array(sorted(no_grad))
    stat, idx, val = grad_summary(depth.grad.numpy())
    out["gstat/depth_features"], out["gidx/depth_features"], out["gval/depth_features"] = stat, idx, val
    stat, idx, val = grad_summary(d2d.grad.numpy())
    out["gstat/roi_depth"], out["gidx/roi_depth"], out["gval/roi_depth"] = stat, idx, val
    bn = pred.pos_embed[0]
    out["running_mean"], out["running_var"] = bn.running_mean.numpy().copy(), bn.running_var.numpy().copy()
    out["num_batches_tracked"] = np.array(int(bn.num_batches_tracked))
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: wrote {path} ({os.path.getsize(path) / 1024:.0f} KB) rel_loss={out['rel_loss']:.6f} "
          f"no_grad={no_grad}")


def run_meet_train_case(ns, name, c):
    """One training step of the UNMODIFIED reference VETOPredictor_MEET (train() mode, dropout p = 0, EXPERT_GROUP
    False): the group sampling under random.seed(sample_seed), the per-group CE losses, and the gradients of their sum
    (what tools/relation_train_net.py:451 back-propagates)."""
    import random
    import torch
    from tests.cases import case_rel_labels
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    ds = synth.VG if c["dataset"] == "VG" else synth.GQA
    cfg = ref_shim.make_cfg(ns, predictor=c["predictor"], mode=c["mode"], dataset=c["dataset"])
    sd = case_state(c)
    pred = ref_shim.build_predictor(ns, cfg, ds["num_obj"], ds["num_rel"], synth.to_torch_state(sd)).train()
    for m in pred.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    batch = case_batch(c)
    bls = ref_shim.make_boxlists(ns, batch, ds["num_obj"])
    fe = ns.make_extractor(cfg, 256, for_relation=True).train()
    samp = ns.make_sampler(cfg)
    pairs = samp.prepare_test_pairs(torch.device("cpu"), bls)
    rel_labels = [torch.from_numpy(l) for l in case_rel_labels(c, [len(p) for p in pairs])]
    feats = [torch.from_numpy(f) for f in batch["feats"]]
    feats.append(torch.zeros(batch["B"], 256, 1, 1))
    depth = torch.from_numpy(batch["depth"]).clone().requires_grad_(True)
    x2d, d2d, _, _ = fe(feats, bls, depth_features=depth)
    d2d.retain_grad()
    random.seed(c["sample_seed"])
    obj_d, rel_d, losses, incre, chosen, custom = pred(bls, pairs, rel_labels, None, roi_features=x2d,
                                                       roi_depth_features=d2d)
    assert obj_d is None and rel_d is None and custom is None
    sum(losses.values()).backward()
    out = {"input_digest": np.array(digest(batch["feats"] + [batch["depth"]] + batch["boxes"] + batch["labels"])),
           "weight_digest": np.array(digest([sd[k] for k in sorted(sd)])),
           "rel_labels": np.concatenate([l.numpy() for l in rel_labels]),
           "pair_counts": np.array([len(p) for p in pairs]),
           "incre_idx_list": np.array(incre),
           "loss_names": np.array(sorted(losses)),
           "losses": np.array([losses[k].item() for k in sorted(losses)], np.float64),
           "expert_dist_len": np.array(len(chosen))}
    for k, rows in enumerate(chosen[0]):
        out[f"chosen/{k}"] = np.array(rows, dtype=np.int64)
    no_grad = []
    for k, p in pred.named_parameters():
        if p.grad is None:
            no_grad.append(k)
            continue
        stat, idx, val = grad_summary(p.grad.numpy())
        out["gstat/" + k], out["gidx/" + k], out["gval/" + k] = stat, idx, val
    out["no_grad"] = np.array(sorted(no_grad))
    stat, idx, val = grad_summary(d2d.grad.numpy())
    out["gstat/roi_depth"], out["gidx/roi_depth"], out["gval/roi_depth"] = stat, idx, val
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: wrote {path} ({os.path.getsize(path) / 1024:.0f} KB) losses={dict(zip(out['loss_names'], out['losses']))} "
          f"chosen sizes={[len(r) for r in chosen[0]]} no_grad={no_grad}")


def run_sample_rates():
    """generate_sample_rate_vector_sep2 / get_current_predicate_idx of the reference for every (dataset, split) the
    drop-in knows, and its group sampling on seeded label vectors (the loop of VETOPredictor_MEET.forward :3940-3969
    needs a predictor instance, so the sampled rows are taken from run_meet_train_case instead)."""
    sys.path.insert(0, ref_shim.REF_ROOT)
    from SHA_GCL_extra.extra_function_utils import (generate_num_stage_vector, generate_sample_rate_vector_sep2,
                                                    get_current_predicate_idx)
    from SHA_GCL_extra.group_chosen_function import get_group_splits
    from veto_b200 import config as vcfg
    out = {}
    for (ds, split), sizes in vcfg.GROUP_SPLITS.items():
        try:
            stage_list, stage_count = get_group_splits(ds, split)
        except (AssertionError, SystemExit):
            continue
        if stage_list is None:
            continue
        assert list(stage_count) == list(sizes), (ds, split)
        out[f"rates/{ds}/{split}"] = np.array(generate_sample_rate_vector_sep2(ds, generate_num_stage_vector(stage_list)))
        out[f"incre/{ds}/{split}"] = np.array(get_current_predicate_idx(stage_list, 0.1, ds)[0])
    # GLOBAL_SETTING.BETA_LOSS: the reference reads a hard-coded absolute path (roi_relation_predictors.py:4059) and
    # calls .cuda(); redirect that one open() to the copy shipped in its repository and neutralise .cuda()
    import builtins
    import torch
    ns = ref_shim.load()
    real_open, real_cuda = builtins.open, torch.Tensor.cuda
    def _open(f, *a, **k):
        if isinstance(f, str) and f.endswith("VETO_rebuttal/pred_counts.pkl"):
            f = os.path.join(ref_shim.REF_ROOT, "pred_counts.pkl")
        return real_open(f, *a, **k)
    builtins.open, torch.Tensor.cuda = _open, (lambda self, *a, **k: self)
    try:
        cfg = ref_shim.make_cfg(ns)
        cfg.merge_from_list(["GLOBAL_SETTING.BETA_LOSS", True])
        pred = ref_shim.build_predictor(ns, cfg, 151, 51)
        out["beta_loss_weight"] = pred.criterion_loss_rel.weight.numpy().copy()
    finally:
        builtins.open, torch.Tensor.cuda = real_open, real_cuda
    path = os.path.join(OUT, "meet_sample_rates.npz")
    np.savez_compressed(path, **out)
    print(f"meet_sample_rates: wrote {path}", sorted(out))


def run_relsample_case(ns, name, c):
    """RelationSampling.gtbox_relsample of the UNMODIFIED reference on seeded relation matrices (CPU, torch.manual_seed)."""
    import torch
    cfg = ref_shim.make_cfg(ns)
    cfg.merge_from_list(["MODEL.ROI_RELATION_HEAD.BATCH_SIZE_PER_IMAGE", c["caps"][0],
                         "MODEL.ROI_RELATION_HEAD.POSITIVE_FRACTION", c["caps"][1]])
    samp = ns.make_sampler(cfg)
    mats = synth.make_relation_matrices(c["seed"], c["n_boxes"], 51, c["fg_per_image"])
    props, tgts = [], []
    for n, m in zip(c["n_boxes"], mats):
        box = torch.rand(n, 4) * 100
        props.append(ns.BoxList(box.clone(), (416, 320), mode="xyxy"))
        t = ns.BoxList(box.clone(), (416, 320), mode="xyxy")
        t.add_field("relation", torch.from_numpy(m))
        tgts.append(t)
    torch.manual_seed(c["seed"])
    props, rel_labels, rel_idx_pairs, binarys = samp.gtbox_relsample(props, tgts)
    out = {"n_images": np.array(len(mats))}
    for i in range(len(mats)):
        out[f"pairs/{i}"] = rel_idx_pairs[i].numpy()
        out[f"labels/{i}"] = rel_labels[i].numpy()
        out[f"binary/{i}"] = binarys[i].numpy()
        out[f"locating_match/{i}"] = props[i].get_field("locating_match").numpy()
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: wrote {path}", [(len(p), int((l > 0).sum())) for p, l in zip(rel_idx_pairs, rel_labels)])


def run_detect_sample_case(ns, name, c):
    """RelationSampling.detect_relsample of the UNMODIFIED reference on seeded detections / ground truth (CPU; numpy and
    torch seeded)."""
    import torch
    cfg = ref_shim.make_cfg(ns, mode="sgdet")
    cfg.merge_from_list(["MODEL.ROI_RELATION_HEAD.BATCH_SIZE_PER_IMAGE", c["caps"][0],
                         "MODEL.ROI_RELATION_HEAD.POSITIVE_FRACTION", c["caps"][1],
                         "MODEL.ROI_RELATION_HEAD.REQUIRE_BOX_OVERLAP", c["require_overlap"]])
    samp = ns.make_sampler(cfg)
    imgs = synth.make_detect_case(c["seed"], c["n_tgt"])
    props, tgts = [], []
    for im in imgs:
        p = ns.BoxList(torch.from_numpy(im["prp_boxes"]), im["size"], mode="xyxy")
        p.add_field("labels", torch.from_numpy(im["prp_labels"]))
        p.add_field("pred_scores", torch.from_numpy(im["prp_scores"]))
        t = ns.BoxList(torch.from_numpy(im["tgt_boxes"]), im["size"], mode="xyxy")
        t.add_field("labels", torch.from_numpy(im["tgt_labels"]))
        t.add_field("relation", torch.from_numpy(im["relation"]))
        props.append(p)
        tgts.append(t)
    np.random.seed(c["seed"])
    torch.manual_seed(c["seed"])
    props, rel_labels, rel_labels_all, rel_idx_pairs, binarys = samp.detect_relsample(props, tgts)
    out = {"n_images": np.array(len(imgs))}
    for i in range(len(imgs)):
        out[f"pairs/{i}"] = rel_idx_pairs[i].numpy()
        out[f"labels/{i}"] = rel_labels[i].numpy()
        out[f"binary/{i}"] = binarys[i].numpy()
        out[f"locating_match/{i}"] = props[i].get_field("locating_match").numpy()
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: wrote {path}", [(len(p), int((l > 0).sum())) for p, l in zip(rel_idx_pairs, rel_labels)])


def run_eval_case(ns, name, c):
    """SGRecall.calculate_recall of the UNMODIFIED reference (sgg_eval.py:138-186) on seeded predictions / ground truth."""
    sys.path.insert(0, ref_shim.REF_ROOT)
    from pysgg.data.datasets.evaluation.vg.sgg_eval import (SGMeanRecall, SGNGMeanRecall, SGNoGraphConstraintRecall,
                                                             SGPairAccuracy, SGRecall, SGZeroShotRecall)
    imgs = synth.make_eval_case(c["seed"], c["n_objs"], c["n_gt_rels"], c["n_pred_rels"])
    result_dict = {}
    ev = SGRecall(result_dict)
    ev.register_container("sgdet")
    mr = SGMeanRecall(result_dict, 51, ["__background__"] + [f"rel{i}" for i in range(1, 51)])
    mr.register_container("sgdet")
    ng = SGNoGraphConstraintRecall(result_dict)
    ng.register_container("sgdet")
    zs = SGZeroShotRecall(result_dict)
    zs.register_container("sgdet")
    ngm = SGNGMeanRecall(result_dict, 51, ["__background__"] + [f"rel{i}" for i in range(1, 51)])
    ngm.register_container("sgdet")
    pa = SGPairAccuracy(result_dict)
    pa.register_container("sgcls")
    # "zero-shot" triplets: (subject class, object class, predicate) of every third ground-truth relation of the case
    zs_list = []
    for im in imgs:
        r = im["relation_tuple"][::3]
        zs_list.append(np.column_stack((im["labels"][r[:, 0]], im["labels"][r[:, 1]], r[:, 2])))
    zs_trip = np.unique(np.concatenate(zs_list), axis=0)
    out = {"n_images": np.array(len(imgs)), "zeroshot_triplets": zs_trip}
    for i, im in enumerate(imgs):
        local = dict(pred_rel_inds=im["rel_pair_idxs"], rel_scores=im["pred_rel_scores"], gt_rels=im["relation_tuple"],
                     gt_classes=im["labels"], gt_boxes=im["boxes"], pred_classes=im["pred_labels"],
                     pred_boxes=im["pred_boxes"], obj_scores=np.ones(len(im["labels"]), np.float32))
        local = ev.calculate_recall({"iou_thres": 0.5}, local, "sgdet")
        mr.collect_mean_recall_items({"iou_thres": 0.5}, local, "sgdet")
        zs.prepare_zeroshot({"zeroshot_triplet": zs_trip}, local)
        zs.calculate_recall({"iou_thres": 0.5}, local, "sgdet")
        # pair accuracy (SGCls-style: boxes from the ground truth are not required by the metric itself)
        loc2 = dict(local)
        pa.prepare_gtpair(loc2)
        keep = pa.pred_pair_in_gt
        loc2["pred_to_gt"] = local["pred_to_gt"]
        pa.calculate_recall({"iou_thres": 0.5}, loc2, "sgcls")
        local["obj_scores"] = im["pred_scores"]
        ng.calculate_recall({"iou_thres": 0.5}, local, "sgdet")
        ngm.collect_mean_recall_items({"iou_thres": 0.5}, local, "sgdet")
        nfirst = np.full(len(im["relation_tuple"]), 2 ** 31 - 1, np.int64)
        for p, gs in enumerate(local["nogc_pred_to_gt"]):
            for g in gs:
                nfirst[g] = min(nfirst[g], p)
        out[f"nogc_first_match/{i}"] = nfirst
        p2g = local["pred_to_gt"]
        first = np.full(len(im["relation_tuple"]), 2 ** 31 - 1, np.int64)
        for p, gs in enumerate(p2g):
            for g in gs:
                first[g] = min(first[g], p)
        out[f"first_match/{i}"] = first
        out[f"pred_hits/{i}"] = np.array([len(g) for g in p2g], np.int64)
    mr.calculate_mean_recall("sgdet")
    ngm.calculate_mean_recall("sgdet")
    for k in (20, 50, 100):
        out[f"mean_recall/{k}"] = np.array(result_dict["sgdet_mean_recall"][k], np.float64)
        out[f"ng_mean_recall/{k}"] = np.array(result_dict["sgdet_ng_mean_recall"][k], np.float64)
        out[f"mean_recall_list/{k}"] = np.array(result_dict["sgdet_mean_recall_list"][k], np.float64)
        out[f"recall/{k}"] = np.array(result_dict["sgdet_recall"][k], np.float64)
        out[f"recall_nogc/{k}"] = np.array(result_dict["sgdet_recall_nogc"][k], np.float64)
        out[f"zeroshot_recall/{k}"] = np.array(result_dict["sgdet_zeroshot_recall"][k], np.float64)
        out[f"accuracy_hit/{k}"] = np.array(result_dict["sgcls_accuracy_hit"][k], np.float64)
        out[f"accuracy_count/{k}"] = np.array(result_dict["sgcls_accuracy_count"][k], np.float64)
        per = {}
        for d in result_dict["sgdet_recall_per_rel"][k]:
            for r, (h, n) in d.items():
                e = per.setdefault(int(r), [0, 0])
                e[0] += int(h)
                e[1] += int(n)
        out[f"per_rel/{k}"] = np.array([[r, h, n] for r, (h, n) in sorted(per.items())], np.int64)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: wrote {path}", {k: out[f'recall/{k}'].round(3).tolist() for k in (20, 50, 100)})


def run_depth_backbone(name="depth_backbone", batch=2, height=70, width=101):
    """The UNMODIFIED reference depth backbone (backbone.py:83-93 build_resnet18_depth -> ResNetDepth) on a synthetic
    depth batch: eval-mode output, then one training-mode forward + backward from a fixed output gradient (parameter
    gradients summarised, updated running statistics).  Weights come from oracle.depth_port.synth_state (numpy RNG), so
    the fixture needs to hold no weights."""
    import importlib
    import torch
    import torchvision.models.resnet as tvr
    from oracle import depth_port
    if not hasattr(tvr, "model_urls"):          # removed from torchvision 0.13+; resnet_depth.py:5 imports the name
        tvr.model_urls = {}
    bb = importlib.import_module("pysgg.modeling.backbone.backbone")
    torch.manual_seed(0)
    model = bb.build_resnet18_depth(None, True)
    keys = list(model.state_dict().keys())
    sd = depth_port.synth_state(0)
    model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=True)
    depth = depth_port.synth_depth(batch, height, width)
    out = {"keys": np.array(keys), "out_channels": np.array(model.out_channels), "shape": np.array(depth.shape),
           "input_digest": np.array(digest([depth])), "weight_digest": np.array(digest([sd[k] for k in sorted(sd)]))}
    model.eval()
    with torch.no_grad():
        out["eval_out"] = model(torch.from_numpy(depth)).numpy()
    model.train()
    y = model(torch.from_numpy(depth))
    g = np.random.RandomState(5).standard_normal(tuple(y.shape)).astype(np.float32)
    y.backward(torch.from_numpy(g))
    out["train_out"] = y.detach().numpy()
    out["grad_out_digest"] = np.array(digest([g]))
    for k, p_ in model.named_parameters():
        stat, idx, val = grad_summary(p_.grad.numpy())
        out["gstat/" + k], out["gidx/" + k], out["gval/" + k] = stat, idx, val
    for k, b in model.named_buffers():
        out["buf/" + k] = b.numpy().copy()
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: wrote {path} ({os.path.getsize(path) / 1024:.0f} KB) out {tuple(y.shape)}")


def main():
    from tests.cases import (DETECT_SAMPLE_CASES, EVAL_CASES, FULL_CASES, FULL_TRAIN_CASES, MEET_TRAIN_CASES,
                             RELSAMPLE_CASES, TRAIN_CASES)
    ns = ref_shim.load()
    only = sys.argv[1:]
    for name, c in list(CASES.items()) + list(FULL_CASES.items()):
        if only and name not in only:
            continue
        run_case(ns, name, c)
    for name, c in list(TRAIN_CASES.items()) + list(FULL_TRAIN_CASES.items()):
        if only and name not in only:
            continue
        run_train_case(ns, name, c)
    for name, c in MEET_TRAIN_CASES.items():
        if only and name not in only:
            continue
        run_meet_train_case(ns, name, c)
    for name, c in RELSAMPLE_CASES.items():
        if only and name not in only:
            continue
        run_relsample_case(ns, name, c)
    for name, c in DETECT_SAMPLE_CASES.items():
        if only and name not in only:
            continue
        run_detect_sample_case(ns, name, c)
    for name, c in EVAL_CASES.items():
        if only and name not in only:
            continue
        run_eval_case(ns, name, c)
    if not only or "meet_sample_rates" in only:
        run_sample_rates()
    if not only or "depth_backbone" in only:
        run_depth_backbone()


if __name__ == "__main__":
    main()
