"""GPU parity tests of the training step (pytest -m gpu): VETOPredictor.forward in train() mode + loss.backward()
through the reference-shaped API (VETOFeatureExtractor -> VETOPredictor -> add_losses['rel_loss'].backward()),
compared with the gradient oracle (oracle/torch_port.train_step, pinned against the unmodified reference by
tests/test_oracle.py) and directly with the fixtures the reference's own training step produced
(tests/golden/train_*.npz).

Bar: the loss within 1e-5 relative; every gradient tensor within 1e-3 of its own max magnitude (max |diff| /
max |ref|, the metric north_star states for logits) in fp32 mode and within the stated tensor-core tolerance 2e-3 in
bf16x3 mode (measured: <= 3.3e-4 without dropout, <= 1.03e-3 with the reference's dropout rates — the worst tensor
is the 4-element BatchNorm bias gradient, a cancelling sum over boxes); the single-pass bf16
mode (8 mantissa bits per operand through 24 chained GEMMs and their transposes) has the stated tolerance 0.25
and exists for throughput experiments, not for parity.
"""
import os

import numpy as np
import pytest
import torch

from tests import harness as H
from tests.cases import FULL_TRAIN_CASES, TRAIN_CASES, load_golden
from tests.train_util import check_against_golden, grad_error, oracle_train_step, train_case_inputs
from veto_b200 import config as vcfg
from veto_b200 import ops, registry, synth

pytestmark = pytest.mark.gpu

DEV = torch.device("cuda:0")
GRAD_TOL = {"fp32": 1e-3, "bf16x3": 2e-3, "bf16": 0.25}
LOSS_TOL = {"fp32": 1e-5, "bf16x3": 1e-5, "bf16": 5e-3}


def _set_dropout(pred, p_pos, p_emb, p_attn):
    pred.pos_embed[3].p = p_pos
    tr = pred.fusion_transformer.transformer
    tr.pos_drop.p = p_emb
    for layer in tr.layers:
        layer[0].fn.to_out[1].p = p_attn


def _run_train(c, precision, batch, sd, pairs, rel_labels, dropout=(0.0, 0.0, 0.0), seed=None):
    """One training step through the public API; returns loss, gradients (numpy) and the module."""
    cfg = H.make_cfg(predictor=c["predictor"], mode=c["mode"], dataset=c["dataset"], precision=precision)
    num_obj = vcfg.num_classes(cfg)[0]
    bls = H.boxlists(batch, DEV, num_obj)
    feats, depth = H.device_features(batch, DEV)
    depth.requires_grad_(True)
    fe = registry.make_roi_box_feature_extractor(cfg, 256, for_relation=True).to(DEV).train()
    pred = registry.make_roi_relation_predictor(cfg, 512)
    pred.load_state_dict(synth.to_torch_state(sd), strict=True)
    pred = pred.to(DEV).train()
    _set_dropout(pred, *dropout)
    if seed is not None:
        torch.manual_seed(seed)
    x2d, d2d, _, _ = fe(feats, bls, depth_features=depth)
    d2d.retain_grad()
    out = pred(bls, [torch.from_numpy(p).to(DEV) for p in pairs], [torch.from_numpy(l).to(DEV) for l in rel_labels], None,
               roi_features=x2d, roi_depth_features=d2d)
    assert out[0] is None and out[1] is None and out[3] is None
    losses = out[2]
    losses["rel_loss"].backward()
    torch.cuda.synchronize()
    grads = {k: H.np_(p.grad) for k, p in pred.named_parameters() if p.grad is not None}
    no_grad = sorted(k for k, p in pred.named_parameters() if p.grad is None)
    return dict(loss=float(losses["rel_loss"].detach()), losses=losses, grads=grads, no_grad=no_grad, g_roi_depth=H.np_(d2d.grad),
                g_depth_map=H.np_(depth.grad), pred=pred)


def _report(name, rows):
    os.makedirs("gpurun_out", exist_ok=True)
    with open(os.path.join("gpurun_out", "train_diag.txt"), "a") as f:
        f.write(f"== {name}\n")
        for k, e in rows:
            f.write(f"  {e:10.3e}  {k}\n")


@pytest.mark.parametrize("precision", ["fp32", "bf16x3", "bf16"])
@pytest.mark.parametrize("name", list(TRAIN_CASES))
def test_train_step_matches_oracle_and_reference(name, precision):
    c = TRAIN_CASES[name]
    g = load_golden(name)
    batch, sd, pairs, labels = train_case_inputs(c)
    ref = oracle_train_step(c, batch, sd, pairs, labels)
    mine = _run_train(c, precision, batch, sd, pairs, labels)
    tol = GRAD_TOL[precision]
    rows = [("loss", abs(mine["loss"] - ref["loss"]) / abs(ref["loss"]))]
    for k, gr in ref["grads"].items():
        assert k in mine["grads"], f"no gradient for {k}"
        rows.append((k, grad_error(mine["grads"][k], gr)))
    rows.append(("roi_depth_features", grad_error(mine["g_roi_depth"], ref["g_roi_depth"])))
    _report(f"{name} {precision}", rows)
    bad = [(k, e) for k, e in rows[1:] if not e <= tol]
    assert not bad, f"gradients outside {tol}: {bad[:8]}"
    assert rows[0][1] <= LOSS_TOL[precision], f"loss {mine['loss']} vs {ref['loss']}"
    # the parameters the reference leaves without a gradient stay without one here
    assert mine["no_grad"] == sorted(str(k) for k in g["no_grad"])
    # ... and directly against what the unmodified reference produced (norm + samples of every gradient,
    # including the depth feature map behind the ROIAlign backward)
    grads = dict(mine["grads"])
    grads["roi_depth"] = mine["g_roi_depth"]
    grads["depth_features"] = mine["g_depth_map"]
    check_against_golden(grads, g, 2 * tol)
    assert abs(mine["loss"] - float(g["rel_loss"])) <= 2 * LOSS_TOL[precision] * abs(float(g["rel_loss"]))
    if "obj_loss" in g.files:
        assert abs(float(mine["losses"]["obj_loss"]) - float(g["obj_loss"])) <= 1e-5 * abs(float(g["obj_loss"]))
    bn = mine["pred"].pos_embed[0]
    assert np.allclose(H.np_(bn.running_mean), g["running_mean"], rtol=1e-5)
    assert np.allclose(H.np_(bn.running_var), g["running_var"], rtol=1e-5)
    assert int(bn.num_batches_tracked) == int(g["num_batches_tracked"])


@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
def test_train_step_full_size_matches_reference(precision):
    """BASELINE.json configs[1] at its full size — IMS_PER_BATCH 12 x 20 GT boxes = 4560 pairs, the size the training
    throughput is quoted on (M = 86 640 token rows: the mixed-width weight-gradient tiles, the split-K decisions and the
    full tile schedule) — against the loss and the gradient summaries (norm + 64 samples of every parameter gradient, of
    the ROI depth features and of the depth feature map behind the ROIAlign backward) the unmodified reference produced
    (tests/golden/train_predcls_full.npz)."""
    name = "train_predcls_full"
    c = FULL_TRAIN_CASES[name]
    g = load_golden(name)
    batch, sd, pairs, labels = train_case_inputs(c)
    assert sum(len(p) for p in pairs) == 4560 and np.array_equal(np.concatenate(labels), g["rel_labels"])
    mine = _run_train(c, precision, batch, sd, pairs, labels)
    tol = GRAD_TOL[precision]
    assert abs(mine["loss"] - float(g["rel_loss"])) <= 2 * LOSS_TOL[precision] * abs(float(g["rel_loss"]))
    assert mine["no_grad"] == sorted(str(k) for k in g["no_grad"])
    grads = dict(mine["grads"])
    grads["roi_depth"] = mine["g_roi_depth"]
    grads["depth_features"] = mine["g_depth_map"]
    worst = check_against_golden(grads, g, 2 * tol)
    _report(f"{name} {precision}", sorted(worst.items(), key=lambda kv: -kv[1])[:12])
    bn = mine["pred"].pos_embed[0]
    assert np.allclose(H.np_(bn.running_mean), g["running_mean"], rtol=1e-5)
    assert np.allclose(H.np_(bn.running_var), g["running_var"], rtol=1e-5)


@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
def test_train_step_with_dropout_matches_oracle(precision):
    """Dropout active (the reference's p = 0.1 / 0.35 / 0.35): the library's counter-based masks are regenerated on
    the host (ops.dropout_keep_mask) and fed to the oracle as explicit keep-scale factors."""
    c = TRAIN_CASES["train_predcls"]
    batch, sd, pairs, labels = train_case_inputs(c)
    p_pos, p_emb, p_attn = 0.1, 0.35, 0.35
    seed = 1234
    torch.manual_seed(seed)
    lib_seed = int(torch.randint(0, 2 ** 62, (1,)).item())
    N, R = sum(batch["n_boxes"]), sum(len(p) for p in pairs)
    mk = lambda sub, shape, p: torch.from_numpy(
        (ops.dropout_keep_mask(lib_seed, sub, int(np.prod(shape)), p).astype(np.float32) / (1.0 - p)).reshape(shape))
    drop = {"pos": mk(1, (N, 128), p_pos), "emb": mk(2, (R, 19, 576), p_emb),
            "attn": [mk(16 + l, (R, 19, 576), p_attn) for l in range(6)]}
    frac = float((drop["emb"] > 0).float().mean())
    assert abs(frac - (1 - p_emb)) < 0.01
    ref = oracle_train_step(c, batch, sd, pairs, labels, drop=drop)
    mine = _run_train(c, precision, batch, sd, pairs, labels, dropout=(p_pos, p_emb, p_attn), seed=seed)
    tol = GRAD_TOL[precision]
    rows = [("loss", abs(mine["loss"] - ref["loss"]) / abs(ref["loss"]))]
    for k, gr in ref["grads"].items():
        rows.append((k, grad_error(mine["grads"][k], gr)))
    rows.append(("roi_depth_features", grad_error(mine["g_roi_depth"], ref["g_roi_depth"])))
    _report(f"dropout {precision}", rows)
    assert rows[0][1] <= LOSS_TOL[precision]
    bad = [(k, e) for k, e in rows[1:] if not e <= tol]
    assert not bad, f"gradients outside {tol}: {bad[:8]}"


def test_train_step_is_bitwise_reproducible():
    c = TRAIN_CASES["train_predcls"]
    batch, sd, pairs, labels = train_case_inputs(c)
    a = _run_train(c, "bf16x3", batch, sd, pairs, labels, dropout=(0.1, 0.35, 0.35), seed=5)
    b = _run_train(c, "bf16x3", batch, sd, pairs, labels, dropout=(0.1, 0.35, 0.35), seed=5)
    assert a["loss"] == b["loss"]
    for k in a["grads"]:
        assert np.array_equal(a["grads"][k], b["grads"][k]), k
    d = _run_train(c, "bf16x3", batch, sd, pairs, labels, dropout=(0.1, 0.35, 0.35), seed=6)
    assert d["loss"] != a["loss"]  # another seed, another mask


def test_train_step_lowers_the_loss():
    """A few Adam steps on one batch through the public API: parameters move, the loss falls."""
    c = TRAIN_CASES["train_predcls"]
    batch, sd, pairs, labels = train_case_inputs(c)
    cfg = H.make_cfg(precision="bf16x3")
    bls = H.boxlists(batch, DEV, 151)
    feats, depth = H.device_features(batch, DEV)
    fe = registry.make_roi_box_feature_extractor(cfg, 256, for_relation=True).to(DEV).train()
    pred = registry.make_roi_relation_predictor(cfg, 512)
    pred.load_state_dict(synth.to_torch_state(sd), strict=True)
    pred = pred.to(DEV).train()
    _set_dropout(pred, 0.0, 0.0, 0.0)
    opt = torch.optim.Adam([p for p in pred.parameters() if p.requires_grad], lr=1e-4)
    pr = [torch.from_numpy(p).to(DEV) for p in pairs]
    lb = [torch.from_numpy(l).to(DEV) for l in labels]
    hist = []
    for _ in range(6):
        opt.zero_grad(set_to_none=True)
        x2d, d2d, _, _ = fe(feats, bls, depth_features=depth)
        loss = pred(bls, pr, lb, None, roi_features=x2d, roi_depth_features=d2d)[2]["rel_loss"]
        loss.backward()
        opt.step()
        hist.append(float(loss.detach()))
    assert hist[-1] < hist[0], hist


def _run_meet_train(c, precision, batch, sd, pairs, rel_labels, sample_seed, expert_group=False, dropout=(0.0, 0.0, 0.0)):
    import random
    cfg = H.make_cfg(predictor=c["predictor"], mode=c["mode"], dataset=c["dataset"], precision=precision)
    cfg.ENSEMBLE_LEARNING.EXPERT_GROUP = expert_group
    num_obj = vcfg.num_classes(cfg)[0]
    bls = H.boxlists(batch, DEV, num_obj)
    feats, depth = H.device_features(batch, DEV)
    depth.requires_grad_(True)
    fe = registry.make_roi_box_feature_extractor(cfg, 256, for_relation=True).to(DEV).train()
    pred = registry.make_roi_relation_predictor(cfg, 512)
    pred.load_state_dict(synth.to_torch_state(sd), strict=True)
    pred = pred.to(DEV).train()
    pred.reference_draws = True     # replay the reference's `random` stream: the fixture's sampled pairs
    _set_dropout(pred.model, *dropout)
    x2d, d2d, _, _ = fe(feats, bls, depth_features=depth)
    d2d.retain_grad()
    random.seed(sample_seed)
    out = pred(bls, [torch.from_numpy(p).to(DEV) for p in pairs], [torch.from_numpy(l).to(DEV) for l in rel_labels], None,
               roi_features=x2d, roi_depth_features=d2d)
    assert out[0] is None and out[1] is None and out[5] is None
    losses = out[2]
    sum(losses.values()).backward()           # tools/relation_train_net.py:451
    torch.cuda.synchronize()
    grads = {k: H.np_(p.grad) for k, p in pred.named_parameters() if p.grad is not None}
    no_grad = sorted(k for k, p in pred.named_parameters() if p.grad is None)
    return dict(losses={k: float(v.detach()) for k, v in losses.items()}, grads=grads, no_grad=no_grad,
                g_roi_depth=H.np_(d2d.grad), g_depth_map=H.np_(depth.grad), incre=out[3], chosen=out[4], pred=pred)


@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
def test_meet_train_step_matches_oracle_and_reference(precision):
    """VETOPredictor_MEET in train() mode: the group sampling under the reference's `random`
    stream picks the reference's pairs, the five 'group_k_CE_loss' values and the gradients of their sum match the
    gradient oracle (torch_port.train_step_meet) and the fixture of the unmodified reference."""
    from tests.cases import MEET_TRAIN_CASES
    from tests.train_util import meet_train_case_inputs, oracle_meet_train_step
    from veto_b200 import meet_sampling as MS
    name = "train_meet_vg"
    c = MEET_TRAIN_CASES[name]
    g = load_golden(name)
    batch, sd, pairs, labels = meet_train_case_inputs(c)
    mine = _run_meet_train(c, precision, batch, sd, pairs, labels, c["sample_seed"])
    chosen = mine["chosen"][0]
    assert len(mine["chosen"]) == int(g["expert_dist_len"])
    for k, rows in enumerate(chosen):
        assert np.array_equal(np.array(rows, dtype=np.int64), g[f"chosen/{k}"]), f"group {k}: sampled pairs differ"
    assert list(mine["incre"]) == list(g["incre_idx_list"])
    names = [str(n) for n in g["loss_names"]]
    assert sorted(mine["losses"]) == names
    got = np.array([mine["losses"][n] for n in names])
    assert np.abs(got - g["losses"]).max() <= 4 * LOSS_TOL[precision] * np.abs(g["losses"]).max(), (got, g["losses"])
    table = MS.group_local_labels(np.concatenate(labels).tolist(), chosen, list(g["incre_idx_list"]))
    ref = oracle_meet_train_step(c, batch, sd, pairs, table)
    tol = GRAD_TOL[precision]
    rows = [("losses", float(np.abs(got - ref["losses"]).max() / np.abs(ref["losses"]).max()))]
    for k, gr in ref["grads"].items():
        assert k in mine["grads"], f"no gradient for {k}"
        rows.append((k, grad_error(mine["grads"][k], gr)))
    rows.append(("roi_depth_features", grad_error(mine["g_roi_depth"], ref["g_roi_depth"])))
    _report(f"{name} {precision}", rows)
    bad = [(k, e) for k, e in rows[1:] if not e <= tol]
    assert not bad, f"gradients outside {tol}: {bad[:8]}"
    # each group loss averages only the 11..41 pairs sampled into it (rel_loss averages all of them): 3x the rel_loss bar
    assert rows[0][1] <= 3 * LOSS_TOL[precision]
    assert mine["no_grad"] == sorted(str(k) for k in g["no_grad"])
    grads = dict(mine["grads"])
    grads["roi_depth"] = mine["g_roi_depth"]
    check_against_golden(grads, g, 2 * tol)


@pytest.mark.parametrize("zero_mode", ["rand_insert", "rand_choose", "all_include"])
def test_meet_group_sampling_kernel(zero_mode):
    """veto_meet_group_labels (roi_relation_predictors.py:3940-3969 + 3812-3821 on the device):
    (i) with the draws of Python's `random` injected in the reference's order the [G, R] table is bit-identical to the
    host restatement of the two loops (itself pinned on the reference's sampled rows, tests/test_oracle.py);
    (ii) with its own counter-based draws every (head, predicate) keep rate matches the sample-rate table within
    binomial bounds, backgrounds spread uniformly over the heads; the same seed reproduces the table, another one
    does not."""
    import random
    from veto_b200 import meet_sampling as MS
    from veto_b200.predictor import incre_idx_list
    for ds, num_rel in (("VG", 51), ("GQA", 101)):
        sizes = synth.GROUP_SPLITS[(ds, "divide4")]
        G = len(sizes)
        incre = incre_idx_list(sizes, num_rel)
        rates = MS.sample_rate_matrix(ds, sizes)
        tables = (torch.tensor(incre, dtype=torch.int32, device=DEV), torch.from_numpy(rates).to(DEV),
                  torch.from_numpy(MS.local_label_table(incre, G)).to(DEV))
        rng = np.random.default_rng(5)
        labels = rng.integers(0, num_rel, size=20000)
        labels[rng.random(20000) < 0.3] = 0
        # (i) injected reference stream
        random.seed(99)
        u, heads = MS.reference_draws(labels.tolist(), G, zero_mode)
        got = H.np_(ops.meet_group_labels(torch.from_numpy(labels).to(DEV), *tables, zero_mode, seed=1,
                                          draws=torch.from_numpy(u), bg_heads=torch.from_numpy(heads)))
        random.seed(99)
        chosen = MS.group_sampling(labels.tolist(), incre, rates, G, zero_mode)
        assert np.array_equal(got, MS.group_local_labels(labels.tolist(), chosen, incre))
        # (ii) own draws
        R = 400000
        labels = rng.integers(0, num_rel, size=R)
        t1 = H.np_(ops.meet_group_labels(torch.from_numpy(labels).to(DEV), *tables, zero_mode, seed=1234))
        t2 = H.np_(ops.meet_group_labels(torch.from_numpy(labels).to(DEV), *tables, zero_mode, seed=1234))
        t3 = H.np_(ops.meet_group_labels(torch.from_numpy(labels).to(DEV), *tables, zero_mode, seed=1235))
        assert np.array_equal(t1, t2) and not np.array_equal(t1, t3)
        local = MS.local_label_table(incre, G)
        kept = t1 >= 0
        assert np.array_equal(t1[kept], local[np.nonzero(kept)[0], labels[np.nonzero(kept)[1]]])   # relabelling
        for p in range(num_rel):
            rows = labels == p
            n = int(rows.sum())
            for k in range(G):
                if p == 0:
                    prob = {"rand_insert": 1.0 / G, "rand_choose": 0.6, "all_include": 1.0}[zero_mode]
                else:   # the pair joins head k iff k + 1 < g_p or u <= max over a >= k + 1 of rates[a - 1][p]
                    prob = 1.0 if k + 1 < incre[p] else float(min(1.0, rates[k:, p].max()))
                freq = kept[k, rows].mean()
                sigma = np.sqrt(max(prob * (1 - prob), 1e-12) / n)
                assert abs(freq - prob) <= 6 * sigma + 1e-12, (ds, p, k, freq, prob)
        if zero_mode == "rand_choose":     # a background pair joins all heads or none
            bg = kept[:, labels == 0]
            assert np.all(bg.all(0) | ~bg.any(0))
        if zero_mode == "rand_insert":     # exactly one head
            assert np.all(kept[:, labels == 0].sum(0) == 1)
        # nested heads for foreground pairs: head k implies every head below it
        fg = kept[:, labels > 0]
        assert np.all(fg[:-1] >= fg[1:])


def test_meet_train_device_sampling_is_seeded_by_random():
    """VETOPredictor_MEET.forward(train) with the default device-side draws: no host copy of the labels, the step is
    reproducible under random.seed, the 5th return value materialises the reference-shaped chosen lists on demand."""
    import random
    from tests.cases import MEET_TRAIN_CASES
    from tests.train_util import meet_train_case_inputs
    c = MEET_TRAIN_CASES["train_meet_vg"]
    batch, sd, pairs, labels = meet_train_case_inputs(c)
    cfg = H.make_cfg(predictor="VETOPredictor_MEET", mode="predcls", dataset="VG", precision="bf16x3")
    bls = H.boxlists(batch, DEV, 151)
    feats, depth = H.device_features(batch, DEV)
    fe = registry.make_roi_box_feature_extractor(cfg, 256, for_relation=True).to(DEV).train()
    pred = registry.make_roi_relation_predictor(cfg, 512)
    pred.load_state_dict(synth.to_torch_state(sd), strict=True)
    pred = pred.to(DEV).train()
    assert pred.reference_draws is False
    x2d, d2d, _, _ = fe(feats, bls, depth_features=depth)
    runs = []
    for seed in (7, 7, 8):
        random.seed(seed)
        torch.manual_seed(3)            # the dropout masks of the step
        out = pred(bls, [torch.from_numpy(p).to(DEV) for p in pairs], [torch.from_numpy(l).to(DEV) for l in labels], None,
                   roi_features=x2d, roi_depth_features=d2d)
        runs.append((H.np_(out[4].table), [float(v.detach()) for v in out[2].values()]))
        assert len(out[4]) == sum(len(l) for l in labels) and len(out[4][0]) == 5
        assert out[4][0][0] == np.nonzero(runs[-1][0][0] >= 0)[0].tolist()
    assert np.array_equal(runs[0][0], runs[1][0]) and runs[0][1] == runs[1][1]
    assert not np.array_equal(runs[0][0], runs[2][0])


def test_meet_train_expert_group_and_loss_weights():
    """EXPERT_GROUP True: 3 experts x 5 groups = 15 heads, expert j of group k sees group k's pairs (:3834-3840), so
    with identical weights per expert the 15 losses repeat the 5 group losses.  Unequal loss weights cannot be
    honoured by the fused step and must poison the gradients instead of being silently wrong."""
    from tests.cases import MEET_TRAIN_CASES
    from tests.train_util import meet_train_case_inputs
    c = MEET_TRAIN_CASES["train_meet_vg"]
    batch, sd, pairs, labels = meet_train_case_inputs(c)
    base = _run_meet_train(c, "bf16x3", batch, sd, pairs, labels, c["sample_seed"])
    sizes = synth.GROUP_SPLITS[("VG", "divide4")]
    sd3 = {k: v for k, v in sd.items() if ".rel_out." not in k}
    for j in range(3):
        for k in range(len(sizes)):
            for t in ("weight", "bias"):
                sd3[f"model.rel_out_group.{j}.{k}.{t}"] = sd[f"model.rel_out.{k}.{t}"]
    for k in range(len(sizes)):
        for t in ("weight", "bias"):
            sd3[f"model.rel_out.{k}.{t}"] = sd[f"model.rel_out.{k}.{t}"]
    exp = _run_meet_train(c, "bf16x3", batch, sd3, pairs, labels, c["sample_seed"], expert_group=True)
    assert len(exp["losses"]) == 15
    for k in range(len(sizes)):
        for j in range(3):
            assert abs(exp["losses"]["group_%d%d_CE_loss" % (k, j + 1)] - base["losses"]["group_%d_CE_loss" % k]) <= 1e-6 * abs(base["losses"]["group_%d_CE_loss" % k])
    # trunk gradients: three identical expert sets -> three times the single-set gradient
    key = "model.fusion_transformer.transformer.layers.0.1.fn.net.0.weight"
    assert grad_error(exp["grads"][key], 3.0 * base["grads"][key]) <= 1e-4
    # unequal weights
    cfg = H.make_cfg(predictor="VETOPredictor_MEET", mode="predcls", dataset="VG", precision="bf16x3")
    bls = H.boxlists(batch, DEV, 151)
    feats, depth = H.device_features(batch, DEV)
    fe = registry.make_roi_box_feature_extractor(cfg, 256, for_relation=True).to(DEV).train()
    pred = registry.make_roi_relation_predictor(cfg, 512)
    pred.load_state_dict(synth.to_torch_state(sd), strict=True)
    pred = pred.to(DEV).train()
    x2d, d2d, _, _ = fe(feats, bls, depth_features=depth)
    losses = pred(bls, [torch.from_numpy(p).to(DEV) for p in pairs], [torch.from_numpy(l).to(DEV) for l in labels], None,
                  roi_features=x2d, roi_depth_features=d2d)[2]
    (losses["group_0_CE_loss"] * 2.0 + losses["group_1_CE_loss"]).backward()
    assert bool(torch.isnan(pred.model.location_projection[0].weight.grad).all())


def test_training_branch_errors():
    c = TRAIN_CASES["train_predcls"]
    cfg = H.make_cfg(predictor="VETOPredictor", precision="bf16x3")
    pred = registry.make_roi_relation_predictor(cfg, 512).to(DEV).train()
    with pytest.raises(Exception):            # no proposals / pairs: a training step needs at least one of each
        pred([], [], [], None, roi_features=None, roi_depth_features=None)


@pytest.mark.parametrize("shape", [(1000, 576, 576), (3000, 1728, 576), (2500, 576, 1152), (777, 1152, 576), (300, 200, 72),
                                   (130, 256, 384), (4000, 576, 2048), (513, 64, 136)])
@pytest.mark.parametrize("split_k", [1, 4])
def test_weight_gradient_gemm_mixed_tiles(shape, split_k):
    """gemm_tn2 (dW = dY^T X, both operands in place) with its mixed 256 / 128-wide column tiles: every width
    combination the encoder produces (576 = 256 + 256 + 64-of-128, 1152 = 4 x 256 + 128, 2048 = 8 x 256) plus ragged
    shapes, with and without split-K, against float64."""
    rows, Nw, Kw = shape
    g = torch.Generator(device=DEV).manual_seed(rows + Nw + Kw)
    y = torch.randn(rows, Nw, generator=g, device=DEV)
    x = torch.randn(rows, Kw, generator=g, device=DEV)
    ref = y.double().t() @ x.double()
    for precision, tol in (("bf16x3", 3e-5), ("bf16", 2e-2)):
        out = ops.test_gemm_tn(y, x, precision, split_k, None)
        torch.cuda.synchronize()
        err = float((out.double() - ref).abs().max() / ref.abs().max())
        assert out.shape == (Nw, Kw) and err < tol, (shape, split_k, precision, err)


@pytest.mark.skipif(os.environ.get("VETO_TRAIN_RECOMPUTE") == "1", reason="already the re-computation arm")
def test_train_step_with_recomputation_matches_reference():
    """VETO_TRAIN_RECOMPUTE=1 (LayerNorm / GELU outputs re-computed in the backward pass instead of saved: 16.0 -> 12.1 GB
    at the configs[1] size) must pass the same parity tests — loss and every gradient against the oracle and the
    reference fixtures, small and full size.  The switch is read once per process: the arm runs in a child pytest."""
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, VETO_TRAIN_RECOMPUTE="1", PYTHONPATH=root)
    r = subprocess.run([sys.executable, "-m", "pytest", "tests/test_gpu_train.py", "-q", "-m", "gpu", "-x", "-k",
                        "train_step_matches_oracle_and_reference or full_size or with_dropout or bitwise_reproducible"],
                       cwd=root, env=env, capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " passed" in r.stdout
    # and the footprint really shrinks
    code = ("import ctypes; from veto_b200 import lib as L, ops; cfg = ops.make_config(151, 51, 'bf16x3');"
            "print(L.load().veto_train_workspace_bytes(ctypes.byref(cfg), 240, 4560))")
    sizes = []
    for flag in ("0", "1"):
        rr = subprocess.run([sys.executable, "-c", code], cwd=root, env=dict(env, VETO_TRAIN_RECOMPUTE=flag), capture_output=True,
                            text=True, timeout=300)
        assert rr.returncode == 0, rr.stderr[-2000:]
        sizes.append(int(rr.stdout.strip().splitlines()[-1]))
    assert sizes[1] < 0.8 * sizes[0]
