"""CPU tests: the oracle (oracle/veto_oracle.py + roialign_oracle.c) against the golden fixtures
produced by the unmodified reference (tests/golden/make_golden.py)."""
import hashlib

import numpy as np
import pytest

from oracle import veto_oracle as O
from tests.cases import CASES, case_batch, case_state, load_golden
from tests.util import tie_groups_equal
from veto_b200 import synth

REL_TOL = 2e-5   # oracle (numpy fp32/OpenBLAS) vs reference (torch fp32/MKL): summation order only


def _digest(arrs):
    h = hashlib.sha256()
    for a in arrs:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def _pairs(c, batch):
    return O.prepare_test_pairs(batch["n_boxes"], c.get("max_pairs", 2048),
                                scores=batch.get("pred_scores"), boxes=batch["boxes"],
                                require_overlap=c.get("require_overlap", False) and c["mode"] == "sgdet")


@pytest.fixture(scope="module", params=list(CASES))
def case(request):
    c = CASES[request.param]
    g = load_golden(request.param)
    batch = case_batch(c)
    return request.param, c, g, batch


def test_inputs_reproduce(case):
    name, c, g, batch = case
    assert _digest(batch["feats"] + [batch["depth"]] + batch["boxes"] + batch["labels"]) == str(g["input_digest"])
    sd = case_state(c)
    assert _digest([sd[k] for k in sorted(sd)]) == str(g["weight_digest"])


@pytest.mark.parametrize("name", ["meet_gqa_full", "train_predcls_full"])
def test_full_size_fixtures_reproduce_and_pin_the_gather(name):
    """The BASELINE-size fixtures (configs[3] B = 16 GQA, configs[1] 12 x 20 training step): the seeded inputs / weights
    regenerate to the digests the reference run recorded, pairs are the reference's, and (configs[3]) the C ROIAlign
    oracle reproduces the stored ROI feature channels bit-exactly.  The logits / gradients of these fixtures are compared
    on the GPU only (the reference itself needed minutes for them)."""
    from tests.cases import FULL_CASES, FULL_TRAIN_CASES
    c = FULL_CASES.get(name) or FULL_TRAIN_CASES[name]
    g = load_golden(name)
    batch = case_batch(c)
    assert _digest(batch["feats"] + [batch["depth"]] + batch["boxes"] + batch["labels"]) == str(g["input_digest"])
    sd = case_state(c)
    assert _digest([sd[k] for k in sorted(sd)]) == str(g["weight_digest"])
    pairs = O.prepare_test_pairs(batch["n_boxes"])
    assert [len(p) for p in pairs] == list(g["pair_counts"])
    if "pairs" in g.files:
        assert np.array_equal(np.concatenate(pairs), g["pairs"])
        x2d, d2d = O.pooler_forward(batch["feats"], batch["depth"], batch["boxes"])
        fs = c["feat_stride"]
        assert np.array_equal(x2d[:, ::fs], g["x2d_sub"]) and np.array_equal(d2d[:, ::fs], g["d2d_sub"])


def test_pairs_bit_exact(case):
    name, c, g, batch = case
    pairs = _pairs(c, batch)
    assert [len(p) for p in pairs] == list(g["pair_counts"])
    if "max_pairs" not in c:
        assert np.array_equal(np.concatenate(pairs), g["pairs"])
        return
    # over the cap the order is (score desc) with the reference's unstable sort deciding ties
    off = 0
    for k, p in enumerate(pairs):
        ref = g["pairs"][off:off + len(p)]
        off += len(p)
        s = batch["pred_scores"][k]
        key = s[ref[:, 0]] * s[ref[:, 1]]
        if len(p) == c["max_pairs"]:
            assert np.all(np.diff(key) <= 0)
            assert tie_groups_equal(p, ref, key)
        else:
            assert np.array_equal(p, ref)


def test_gather_bit_exact(case):
    name, c, g, batch = case
    x2d, d2d = O.pooler_forward(batch["feats"], batch["depth"], batch["boxes"])
    assert np.array_equal(x2d[:, ::16], g["x2d_sub"])
    assert np.array_equal(d2d[:, ::16], g["d2d_sub"])
    assert np.array_equal(x2d.sum(axis=(1, 2, 3), dtype=np.float64), g["x2d_sum"])
    assert np.array_equal(d2d.sum(axis=(1, 2, 3), dtype=np.float64), g["d2d_sum"])


def test_all_levels_populated():
    b = case_batch(CASES["cfg1_predcls_vg"], features=False)
    assert set(O.level_map(b["boxes"]).tolist()) == {0, 1, 2, 3}
    assert np.array_equal(O.level_map(b["boxes"]), synth.box_levels(np.concatenate(b["boxes"])))


def test_logits(case):
    name, c, g, batch = case
    sd = case_state(c)
    pairs = _pairs(c, batch)
    if "max_pairs" in c:      # row alignment with the golden logits: use the reference's tie order
        pairs = np.split(g["pairs"], np.cumsum(g["pair_counts"])[:-1])
    x2d, d2d = O.pooler_forward(batch["feats"], batch["depth"], batch["boxes"])
    if c["predictor"].endswith("MEET"):
        gs = synth.GROUP_SPLITS[(c["dataset"], "divide4")]
        out = O.meet_forward(sd, batch, pairs, x2d, d2d, gs, c["mode"])
        ds = synth.VG if c["dataset"] == "VG" else synth.GQA
        assert O.incre_idx_list(gs, ds["num_rel"]) == list(g["incre_idx_list"])
        for k, v in out.items():
            ref = g["logits_" + k]
            assert v.shape == ref.shape
            assert np.abs(v - ref).max() <= REL_TOL * np.abs(ref).max(), k
            assert np.array_equal(v.argmax(1), ref.argmax(1))
        return
    logits = O.predictor_forward(sd, batch, pairs, x2d, d2d, c["mode"])
    ref = g["logits"]
    assert np.abs(logits - ref).max() <= REL_TOL * np.abs(ref).max()
    assert np.array_equal(logits.argmax(1), ref.argmax(1))
    if c["spread"] and len(ref) >= 300:
        assert len(np.unique(ref[:, 1:].argmax(1))) >= 10      # the argmax check is not vacuous
    if "tokens" in g.files:
        tok = O.tokens_only(sd, batch, pairs, x2d, d2d, c["mode"])[g["token_rows"]]
        assert np.abs(tok - g["tokens"]).max() <= REL_TOL * np.abs(g["tokens"]).max()
    if "post_obj_labels" in g.files:
        # SGDet branch of the post-processor (inference.py:414-431): late per-class NMS on the reference's own softmax
        # tile, per-class box regression, triple ranking
        split = np.cumsum([len(p) for p in pairs])[:-1]
        bsplit = np.cumsum(batch["n_boxes"])[:-1]
        res = O.postprocess_sgdet(np.split(ref, split), batch["predict_logits"], pairs, batch["boxes_per_cls"], 0.5,
                                  obj_scores_softmax=np.split(g["post_obj_prob"], bsplit))
        assert np.array_equal(np.concatenate([r["obj_pred"] for r in res]), g["post_obj_labels"])
        assert (g["post_obj_labels"] != np.concatenate(batch["pred_labels"])).sum() >= 2
        assert np.array_equal(np.concatenate([r["obj_scores"] for r in res]), g["post_obj_scores"])
        assert np.array_equal(np.concatenate([r["boxes"] for r in res]), g["post_boxes"])
        s = np.concatenate([r["triple"] for r in res])
        distinct = np.ones(len(s), bool)
        distinct[1:] &= s[1:] != s[:-1]
        distinct[:-1] &= s[:-1] != s[1:]
        assert np.array_equal(np.concatenate([r["pairs"] for r in res])[distinct], g["post_pairs"][distinct])
        assert np.array_equal(np.concatenate([r["labels"] for r in res])[distinct], g["post_labels"][distinct])
        assert np.allclose(np.concatenate([r["probs"] for r in res])[:, 1:].max(1), g["post_scores"], rtol=1e-5)
        # the numpy softmax of the oracle gives the same labels as the reference's torch softmax on this case
        res2 = O.postprocess_sgdet(np.split(ref, split), batch["predict_logits"], pairs, batch["boxes_per_cls"], 0.5)
        assert np.array_equal(np.concatenate([r["obj_pred"] for r in res2]), g["post_obj_labels"])
    elif "post_pairs" in g.files:
        obj_logits, off = [], 0
        for lab in batch["labels"]:
            ol = np.full((len(lab), batch["num_obj"]), -1000.0, np.float32)    # to_onehot, model_kern.py:266-281
            ol[np.arange(len(lab)), lab] = 1000.0
            obj_logits.append(ol)
        split = np.cumsum([len(p) for p in pairs])[:-1]
        res = O.postprocess(np.split(ref, split), obj_logits, pairs)
        # ranking parity is tie-aware: scores must agree, and wherever scores are distinct so must rows
        s = np.concatenate([r["triple_scores"] for r in res])
        assert np.allclose(s, g["post_scores"], rtol=1e-6, atol=0)
        pp = np.concatenate([r["rel_pair_idxs"] for r in res])
        distinct = np.ones(len(s), bool)
        distinct[1:] &= s[1:] != s[:-1]
        distinct[:-1] &= s[:-1] != s[1:]
        assert np.array_equal(pp[distinct], g["post_pairs"][distinct])
        assert np.array_equal(np.concatenate([r["pred_rel_labels"] for r in res])[distinct], g["post_labels"][distinct])


def test_empty_and_degenerate_pairs():
    p = O.prepare_test_pairs([1, 2, 0])
    assert np.array_equal(p[0], [[0, 0]]) and np.array_equal(p[1], [[0, 1], [1, 0]]) and np.array_equal(p[2], [[0, 0]])
    for n in (3, 20, 80):
        q = O.prepare_test_pairs([n], max_pairs=10 ** 9)[0]
        r = np.arange(n * (n - 1))
        i, j = r // (n - 1), r % (n - 1)
        j = j + (j >= i)
        assert np.array_equal(q, np.stack([i, j], 1))          # closed form used by the CUDA kernel


def _ref_lib():
    import ctypes
    import os
    path = os.path.join(os.path.dirname(O.__file__), "_ref", "libref_roialign.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref not built (needs /root/reference: bash oracle/build_ref.sh)")
    import torch  # noqa: F401  (libtorch must be resident before the reference .so is opened)
    lib = ctypes.CDLL(path)
    fp = ctypes.POINTER(ctypes.c_float)
    lib.ref_roi_align_forward.argtypes = [fp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, fp, ctypes.c_int,
                                          ctypes.c_float, ctypes.c_int, ctypes.c_int, ctypes.c_int, fp]
    return lib


@pytest.mark.parametrize("scale,hw", [(0.25, (40, 52)), (0.125, (20, 26)), (0.0625, (10, 13)), (0.03125, (5, 7))])
def test_roialign_c_oracle_equals_compiled_reference(scale, hw):
    """oracle/roialign_oracle.c vs the unmodified reference ROIAlign_cpu.cpp (oracle/_ref): bit-exact,
    including boxes hanging outside the map, degenerate (<1 px) boxes and sampling_ratio<=0."""
    import ctypes
    lib = _ref_lib()
    rng = np.random.default_rng(int(scale * 1e5))
    B, C = 2, 5
    inp = rng.standard_normal((B, C) + hw, dtype=np.float32)
    W, H = hw[1] / scale, hw[0] / scale
    boxes = synth.make_boxes(rng, 12, int(W), int(H), max_side=min(W, H) * 0.9)
    boxes = np.concatenate([boxes, np.array([[-30, -20, 40, 50], [W - 10, H - 10, W + 60, H + 40],
                                             [10, 10, 10.2, 10.1], [W + 50, H + 50, W + 90, H + 90]], np.float32)])
    rois = np.concatenate([rng.integers(0, B, (len(boxes), 1)).astype(np.float32), boxes], 1)
    for sr in (2, 0, 3):
        mine = O.roi_align(inp, rois, scale, 8, 8, sr)
        ref = np.empty_like(mine)
        fp = ctypes.POINTER(ctypes.c_float)
        lib.ref_roi_align_forward(inp.ctypes.data_as(fp), B, C, hw[0], hw[1], rois.ctypes.data_as(fp), len(rois),
                                  scale, 8, 8, sr, ref.ctypes.data_as(fp))
        assert np.array_equal(mine, ref), sr


def test_roialign_backward_matches_torchvision():
    import torch
    import torchvision
    rng = np.random.default_rng(7)
    inp_shape = (2, 3, 10, 13)
    boxes = synth.make_boxes(rng, 9, 208, 160, max_side=140)
    rois = np.concatenate([rng.integers(0, 2, (9, 1)).astype(np.float32), boxes], 1)
    g = rng.standard_normal((9, 3, 8, 8), dtype=np.float32)
    mine = O.roi_align_backward(g, rois, inp_shape, 0.0625, 2)
    ref = torch.ops.torchvision._roi_align_backward(torch.from_numpy(g), torch.from_numpy(rois), 0.0625, 8, 8,
                                                    2, 3, 10, 13, 2, False).numpy()
    assert np.allclose(mine, ref, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("name", ["cfg1_predcls_vg", "ragged_predcls", "sgdet_cap"])
def test_torch_port_matches_reference(name):
    """oracle/torch_port.py (the CPU baseline bench.py times) against the golden outputs of the unmodified reference:
    pairs and ROI gather bit-exact, logits within summation-order noise."""
    import torch
    from oracle import torch_port as TP
    c = CASES[name]
    g = load_golden(name)
    batch = case_batch(c)
    sd = TP.to_torch(case_state(c))
    boxes = [torch.from_numpy(b) for b in batch["boxes"]]
    scores = [torch.from_numpy(s) for s in batch["pred_scores"]] if "pred_scores" in batch else None
    with torch.no_grad():
        pairs = TP.prepare_test_pairs(batch["n_boxes"], c.get("max_pairs", 2048), scores)
        assert [len(p) for p in pairs] == list(g["pair_counts"])
        if "max_pairs" not in c:
            assert np.array_equal(torch.cat(pairs).numpy(), g["pairs"])
        pairs = [torch.from_numpy(p) for p in np.split(g["pairs"], np.cumsum(g["pair_counts"])[:-1])]
        x2d, d2d = TP.pooler_forward([torch.from_numpy(f) for f in batch["feats"]], torch.from_numpy(batch["depth"]), boxes)
        assert np.array_equal(x2d.numpy()[:, ::16], g["x2d_sub"]) and np.array_equal(d2d.numpy()[:, ::16], g["d2d_sub"])
        logits = TP.predictor_forward(sd, boxes, pairs, x2d, d2d, c["mode"],
                                      labels=[torch.from_numpy(l) for l in batch["labels"]],
                                      predict_logits=[torch.from_numpy(l) for l in batch.get("predict_logits", [])] or None)
    ref = g["logits"]
    assert np.abs(logits.numpy() - ref).max() <= REL_TOL * np.abs(ref).max()
    assert np.array_equal(logits.numpy().argmax(1), ref.argmax(1))


@pytest.mark.parametrize("name", ["train_predcls", "train_sgcls"])
def test_torch_port_train_step_matches_reference(name):
    """oracle/torch_port.train_step (the gradient oracle of the GPU training tests) against one training step of the
    unmodified reference (train() mode, dropout p = 0): rel_loss, every parameter gradient, the gradient wrt the
    pooled depth features, and the BatchNorm running statistics."""
    from tests.cases import TRAIN_CASES
    from tests.train_util import check_against_golden, oracle_train_step, train_case_inputs
    c = TRAIN_CASES[name]
    g = load_golden(name)
    batch, sd, pairs, labels = train_case_inputs(c)
    assert _digest(batch["feats"] + [batch["depth"]] + batch["boxes"] + batch["labels"]) == str(g["input_digest"])
    assert _digest([sd[k] for k in sorted(sd)]) == str(g["weight_digest"])
    assert np.array_equal(np.concatenate(labels), g["rel_labels"])
    out = oracle_train_step(c, batch, sd, pairs, labels)
    assert abs(out["loss"] - float(g["rel_loss"])) <= 2e-5 * abs(float(g["rel_loss"]))
    grads = dict(out["grads"])
    grads["roi_depth"] = out["g_roi_depth"]
    worst = check_against_golden(grads, g, 2e-4, skip=("depth_features",))
    assert len(worst) >= 80                      # every trained parameter was compared
    for k in g["no_grad"]:
        assert str(k) not in out["grads"]
    assert np.allclose(out["running_mean"], g["running_mean"], rtol=1e-6)
    assert np.allclose(out["running_var"], g["running_var"], rtol=1e-6)


def test_meet_sample_rates_and_group_index_match_reference():
    """veto_b200.meet_sampling.sample_rate_matrix / predictor.incre_idx_list against the reference's
    generate_sample_rate_vector_sep2 / get_current_predicate_idx for every (dataset, split) (golden: make_golden.py)."""
    from veto_b200 import config as vcfg
    from veto_b200 import meet_sampling as MS
    from veto_b200.predictor import incre_idx_list
    g = load_golden("meet_sample_rates")
    seen = 0
    for (ds, split), sizes in vcfg.GROUP_SPLITS.items():
        key = f"rates/{ds}/{split}"
        if key not in g.files:
            continue
        seen += 1
        ref = g[key]
        mine = MS.sample_rate_matrix(ds, sizes)
        assert mine.shape == ref.shape
        assert np.array_equal(mine, ref), (ds, split, np.abs(mine - ref).max())
        assert incre_idx_list(sizes, ref.shape[1]) == list(g[f"incre/{ds}/{split}"])
    assert seen >= 8


def test_meet_group_sampling_matches_reference():
    """The host-side group sampling + relabelling reproduces the reference's cur_chosen_matrix under the same
    random.seed, and torch_port.train_step_meet reproduces the reference's group losses and gradients."""
    import random
    from tests.cases import MEET_TRAIN_CASES
    from tests.train_util import check_against_golden, meet_train_case_inputs, oracle_meet_train_step
    from veto_b200 import meet_sampling as MS
    name = "train_meet_vg"
    c = MEET_TRAIN_CASES[name]
    g = load_golden(name)
    batch, sd, pairs, labels = meet_train_case_inputs(c)
    assert _digest(batch["feats"] + [batch["depth"]] + batch["boxes"] + batch["labels"]) == str(g["input_digest"])
    assert _digest([sd[k] for k in sorted(sd)]) == str(g["weight_digest"])
    flat = np.concatenate(labels)
    assert np.array_equal(flat, g["rel_labels"])
    sizes = synth.GROUP_SPLITS[("VG", "divide4")]
    incre = list(g["incre_idx_list"])
    random.seed(c["sample_seed"])
    chosen = MS.group_sampling(flat.tolist(), incre, MS.sample_rate_matrix("VG", sizes), len(sizes), "rand_insert")
    for k, rows in enumerate(chosen):
        assert np.array_equal(np.array(rows, dtype=np.int64), g[f"chosen/{k}"]), k
    assert int(g["expert_dist_len"]) == len(flat)
    table = MS.group_local_labels(flat.tolist(), chosen, incre)
    # the relabelling, restated naively (roi_relation_predictors.py:3812-3821)
    for k, rows in enumerate(chosen):
        members = [i for i, x in enumerate(incre) if x == k + 1]
        for r in range(len(flat)):
            if r not in rows:
                assert table[k, r] == -1
            else:
                p = int(flat[r])
                assert table[k, r] == (0 if p == 0 else members.index(p) + 1 if p in members else len(members) + 1)
    out = oracle_meet_train_step(c, batch, sd, pairs, table)
    assert [str(n) for n in g["loss_names"]] == [f"group_{k}_CE_loss" for k in range(len(sizes))]
    assert np.abs(out["losses"] - g["losses"]).max() <= 2e-5 * np.abs(g["losses"]).max()
    grads = dict(out["grads"])
    grads["roi_depth"] = out["g_roi_depth"]
    worst = check_against_golden(grads, g, 2e-4)
    assert len(worst) >= 85
    for k in g["no_grad"]:
        assert str(k) not in out["grads"]


@pytest.mark.parametrize("name", ["relsample_under_caps", "relsample_over_caps"])
def test_gtbox_relsample_candidates_match_reference(name):
    """oracle.gtbox_relsample_candidates (the deterministic part of RelationSampling.gtbox_relsample) against the
    unmodified reference's sampler run on seeded relation matrices: under the caps the foreground rows / labels are
    identical and the background rows are a permutation of the candidates; over the caps the reference's rows are a
    subset of the candidates of the sizes :91-99 prescribe."""
    from tests.cases import RELSAMPLE_CASES
    c, g = RELSAMPLE_CASES[name], load_golden(name)
    mats = synth.make_relation_matrices(c["seed"], c["n_boxes"], 51, c["fg_per_image"])
    batch, num_pos = c["caps"][0], int(c["caps"][0] * c["caps"][1])
    for i, m in enumerate(mats):
        fg, labels, bg, binary = O.gtbox_relsample_candidates(m)
        ref_pairs, ref_labels = g[f"pairs/{i}"].reshape(-1, 2), g[f"labels/{i}"]
        assert np.array_equal(binary, g[f"binary/{i}"])
        n_fg = min(len(fg), num_pos)
        n_bg = min(len(bg), batch - n_fg)
        assert len(ref_pairs) == n_fg + n_bg and int((ref_labels > 0).sum()) == n_fg
        cand = {tuple(p): int(l) for p, l in zip(fg, labels)}
        if len(fg) <= num_pos:
            assert np.array_equal(ref_pairs[:n_fg], fg) and np.array_equal(ref_labels[:n_fg], labels)
        else:
            assert all(cand[tuple(p)] == int(l) for p, l in zip(ref_pairs[:n_fg], ref_labels[:n_fg]))
            assert len({tuple(p) for p in ref_pairs[:n_fg]}) == n_fg
        bgset = {tuple(p) for p in bg}
        got_bg = [tuple(p) for p in ref_pairs[n_fg:]]
        assert len(set(got_bg)) == n_bg and set(got_bg) <= bgset and np.all(ref_labels[n_fg:] == 0)
        if len(bg) <= batch - n_fg:
            assert set(got_bg) == bgset


def test_recall_matching_matches_reference():
    """oracle.compute_pred_matches / recall_at_k against SGRecall.calculate_recall of the unmodified reference."""
    from tests.cases import EVAL_CASES
    c, g = EVAL_CASES["eval_recall"], load_golden("eval_recall")
    imgs = synth.make_eval_case(c["seed"], c["n_objs"], c["n_gt_rels"], c["n_pred_rels"])
    assert int(g["n_images"]) == len(imgs)
    for i, im in enumerate(imgs):
        s, o, p = im["relation_tuple"][:, 0], im["relation_tuple"][:, 1], im["relation_tuple"][:, 2]
        gt_t = np.column_stack((im["labels"][s], p, im["labels"][o]))
        gt_b = np.column_stack((im["boxes"][s], im["boxes"][o]))
        pl = 1 + im["pred_rel_scores"][:, 1:].argmax(1)
        ps, po = im["rel_pair_idxs"][:, 0], im["rel_pair_idxs"][:, 1]
        pr_t = np.column_stack((im["pred_labels"][ps], pl, im["pred_labels"][po]))
        pr_b = np.column_stack((im["pred_boxes"][ps], im["pred_boxes"][po]))
        p2g = O.compute_pred_matches(gt_t, pr_t, gt_b, pr_b, 0.5)
        first = np.full(len(gt_t), 2 ** 31 - 1, np.int64)
        for k, gs in enumerate(p2g):
            for gi in gs:
                first[gi] = min(first[gi], k)
        assert np.array_equal(first, g[f"first_match/{i}"])
        assert np.array_equal(np.array([len(x) for x in p2g]), g[f"pred_hits/{i}"])
        rec = O.recall_at_k(p2g, len(gt_t))
        for k in (20, 50, 100):
            assert rec[k] == g[f"recall/{k}"][i]
    assert g["recall/100"].max() > 0.5 and g["recall/20"].min() < 0.2      # the case discriminates


@pytest.mark.parametrize("name", ["detsample_default", "detsample_tight"])
def test_detect_relsample_candidates_match_reference(name):
    """oracle.detect_relsample_candidates against RelationSampling.detect_relsample of the unmodified reference: the
    deterministic outputs (binary matrix, locating_match) are identical and the reference's sampled rows satisfy every
    structural property the candidate sets imply."""
    from tests.cases import DETECT_SAMPLE_CASES
    from tests.train_util import check_detect_sample
    c, g = DETECT_SAMPLE_CASES[name], load_golden(name)
    imgs = synth.make_detect_case(c["seed"], c["n_tgt"])
    batch, num_pos = c["caps"][0], int(c["caps"][0] * c["caps"][1])
    drew = 0
    for i, im in enumerate(imgs):
        cand = O.detect_relsample_candidates(im["prp_boxes"], im["prp_labels"], im["prp_scores"], im["tgt_boxes"],
                                             im["tgt_labels"], im["relation"], 0.5, c["require_overlap"])
        assert np.array_equal(cand["binary"], g[f"binary/{i}"])
        assert np.array_equal(cand["locating"], g[f"locating_match/{i}"])
        check_detect_sample(cand, g[f"pairs/{i}"], g[f"labels/{i}"], batch, num_pos)
        drew += sum(len(x[3]) > 4 for x in cand["gt"])
    assert drew >= 1                                            # the weighted draw is exercised
