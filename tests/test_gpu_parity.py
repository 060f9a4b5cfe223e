"""GPU parity tests (run on the B200 box: pytest -m gpu).  Everything goes through the C ABI of
libveto_b200.so via the reference-shaped host API, and is compared with the CPU oracle (oracle/) and with the
golden fixtures produced by the unmodified reference (tests/golden/make_golden.py).

Bars (BASELINE.json north_star): rel_pair_idxs, ROI gather and argmax labels bit-exact; relation logits within
1e-3 relative (max |diff| / max |ref|) for the fp32 and bf16x3 modes; the single-pass bf16 tensor-core mode has the
stated tolerance BF16_TOL = 3e-2 and is not required to keep every argmax.
"""
import numpy as np
import pytest
import torch

from oracle import veto_oracle as O
from tests import harness as H
from tests.cases import CASES, FULL_CASES, case_batch, case_state, load_golden
from tests.util import rel_err, tie_groups_equal
from veto_b200 import lib as L
from veto_b200 import ops, synth

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-3     # north_star: "relation logits must match within 1e-3 relative in fp32"
BF16_TOL = 3e-2     # stated tolerance of the single-pass bf16 tensor-core mode
F16_TOL = 5e-3      # stated tolerance of the single-pass fp16 tensor-core mode (measured 2.2e-3, tools/precision_study.py)
# f16c8 (fp16 product + fp8 first-order corrections, 2 MMA units) meets the fp32 bar like bf16x3 (3 units)
TOL = {"fp32": FP32_TOL, "bf16x3": FP32_TOL, "f16c8": FP32_TOL, "bf16": BF16_TOL, "f16": F16_TOL}
DEV = torch.device("cuda:0")


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


# ------------------------------------------------------------------------------------------ kernels
@pytest.mark.parametrize("shape", [(300, 51, 576), (1000, 1152, 200), (77, 128, 128), (130, 60, 576)])
def test_gemm_simt(shape):
    M, N, K = shape
    g = torch.Generator().manual_seed(0)
    a, w, b, r = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g), torch.randn(N, generator=g), \
        torch.randn(M, N, generator=g)
    ref = a.double() @ w.double().T
    assert rel_err(H.np_(ops.test_gemm(_t(a.numpy()), _t(w.numpy()))), ref.numpy()) < 5e-6
    out = ops.test_gemm(_t(a.numpy()), _t(w.numpy()), bias=_t(b.numpy()), residual=_t(r.numpy()), act=2)
    exp = torch.nn.functional.gelu(ref + b.double()) + r.double()
    assert rel_err(H.np_(out), exp.numpy()) < 5e-6


@pytest.mark.parametrize("precision,tol", [("f16c8", 2e-4), ("f16", 2e-3)])
@pytest.mark.parametrize("shape", [(128, 192, 64), (333, 576, 1152), (1000, 1728, 576), (19456, 576, 576), (777, 1152, 576),
                                   (40000, 192, 128)])
def test_gemm_tcgen05_f16c8(shape, precision, tol):
    """The f16c8 product scheme (kind::f16 fp16 product + kind::f8f6f4 e4m3 correction over 2K into the same TMEM
    accumulator) and the single fp16 product, against the fp64 product of the fp32 operands; with bias / GELU, and with
    the f16c8-format OUTPUT (what FF1 hands to FF2) decoded on the host."""
    M, N, K = shape
    g = torch.Generator().manual_seed(1)
    a, w, b = 2.0 * torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) / K ** 0.5, torch.randn(N, generator=g)
    a[::7] *= 40.0                                  # rows of large activations (the residual stream grows with depth)
    w[::5] *= 0.01                                  # rows of small weights (e4m3 subnormals)
    ref = a.double() @ w.double().T
    assert rel_err(H.np_(ops.test_gemm(_t(a.numpy()), _t(w.numpy()), precision=precision)), ref.numpy()) < tol
    out = ops.test_gemm(_t(a.numpy()), _t(w.numpy()), bias=_t(b.numpy()), act=2, precision=precision)
    assert rel_err(H.np_(out), torch.nn.functional.gelu(ref + b.double()).numpy()) < tol
    if N % 64 == 0:
        # operand-format output: hi = fp16, lo = (residual * 256, value / 8) e4m3 bytes per 64-element block
        scratch = torch.empty(4 * (M * K + N * K) + 4 * M * N, dtype=torch.uint8, device=DEV)
        ops.test_gemm(_t(a.numpy()), _t(w.numpy()), bias=_t(b.numpy()), act=0x100, precision=precision, scratch=scratch)
        torch.cuda.synchronize()
        base = 4 * (M * K + N * K)
        hi = scratch[base:base + 2 * M * N].view(torch.float16).reshape(M, N).double().cpu()
        exp = ref + b.double()
        assert rel_err(hi.numpy(), exp.numpy()) < 1e-3          # fp16 rounding of the value
        if precision == "f16c8":
            lo = scratch[base + 2 * M * N:base + 4 * M * N].view(torch.float8_e4m3fn).reshape(M, N // 64, 2, 64).float().cpu()
            res, val = lo[:, :, 0].reshape(M, N).double() / 256.0, lo[:, :, 1].reshape(M, N).double() * 8.0
            assert rel_err((hi + res).numpy(), exp.numpy()) < 1e-4                 # fp16 + e4m3 residual: ~15 bits
            assert rel_err(val.numpy(), exp.numpy()) < 0.07                        # e4m3 copy of the value itself


@pytest.mark.parametrize("precision,tol", [("bf16", 5e-6), ("bf16x3", 3e-5)])
@pytest.mark.parametrize("shape", [(128, 192, 64), (100, 128, 64), (333, 576, 1152), (1000, 1728, 576), (320, 1024, 1024),
                                   (320, 128, 1024), (19456, 576, 576), (700, 64, 576), (1000, 64, 64), (900, 576, 64)])
def test_gemm_tcgen05(shape, precision, tol):
    """tcgen05/TMEM GEMM.  bf16: exact up to fp32 accumulation against the product of the bf16-rounded operands;
    bf16x3: fp32-grade against the fp64 product of the fp32 operands."""
    M, N, K = shape
    g = torch.Generator().manual_seed(1)
    a, w, b = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) / K ** 0.5, torch.randn(N, generator=g)
    if precision == "bf16":
        ref = a.bfloat16().double() @ w.bfloat16().double().T
    else:
        ref = a.double() @ w.double().T
    assert rel_err(H.np_(ops.test_gemm(_t(a.numpy()), _t(w.numpy()), precision=precision)), ref.numpy()) < tol
    out = ops.test_gemm(_t(a.numpy()), _t(w.numpy()), bias=_t(b.numpy()), act=1, precision=precision)
    assert rel_err(H.np_(out), torch.relu(ref + b.double()).numpy()) < tol


def test_layernorm_and_attention():
    g = torch.Generator().manual_seed(2)
    x, w, b = 3 * torch.randn(1000, 576, generator=g) + 0.5, torch.randn(576, generator=g), torch.randn(576, generator=g)
    ref = torch.nn.functional.layer_norm(x.double(), (576,), w.double(), b.double(), 1e-5)
    assert rel_err(H.np_(ops.test_layernorm(_t(x.numpy()), _t(w.numpy()), _t(b.numpy()))), ref.numpy()) < 2e-6
    n_seq = 301
    qkv = torch.randn(n_seq * 19, 1728, generator=g)
    q, k, v = [t.reshape(n_seq, 19, 6, 96).permute(0, 2, 1, 3).double() for t in qkv.chunk(3, -1)]
    att = torch.softmax(q @ k.transpose(-1, -2) * (96 ** -0.5), -1) @ v
    ref = att.permute(0, 2, 1, 3).reshape(n_seq * 19, 576)
    assert rel_err(H.np_(ops.test_attention(_t(qkv.numpy()))), ref.numpy()) < 5e-6


# ------------------------------------------------------------------------------------------ a1 pairs
def _pairs_gpu(c, batch):
    boxes = _t(np.concatenate(batch["boxes"]))
    scores = _t(np.concatenate(batch["pred_scores"])) if "pred_scores" in batch else None
    return ops.enumerate_pairs(batch["n_boxes"], DEV, c.get("max_pairs", 2048), boxes=boxes, scores=scores,
                               require_overlap=c.get("require_overlap", False) and c["mode"] == "sgdet")


def _pairs_oracle(c, batch):
    return O.prepare_test_pairs(batch["n_boxes"], c.get("max_pairs", 2048), scores=batch.get("pred_scores"),
                                boxes=batch["boxes"],
                                require_overlap=c.get("require_overlap", False) and c["mode"] == "sgdet")


@pytest.mark.parametrize("name", list(CASES))
def test_pairs_bit_exact(name):
    c = CASES[name]
    batch = case_batch(c, features=False)
    g = load_golden(name)
    got = [H.np_(p) for p in _pairs_gpu(c, batch)]
    ref = _pairs_oracle(c, batch)
    assert [len(p) for p in got] == list(g["pair_counts"])
    for a, b in zip(got, ref):
        assert a.dtype == np.int64 and np.array_equal(a, b)
    if "max_pairs" not in c:
        assert np.array_equal(np.concatenate(got), g["pairs"])
    else:  # over the cap the reference's unstable sort decides ties: compare tie-aware against the golden rows
        off = 0
        for k, p in enumerate(got):
            gp = g["pairs"][off:off + len(p)]
            off += len(p)
            s = batch["pred_scores"][k]
            assert tie_groups_equal(p, gp, s[gp[:, 0]] * s[gp[:, 1]])


def test_pairs_edge_cases():
    # empty / single-box images give the [[0,0]] placeholder (sampling.py:47-51); large n uses the closed form
    n_boxes = [0, 1, 2, 300, 5]
    got = [H.np_(p) for p in ops.enumerate_pairs(n_boxes, DEV, max_pairs=10 ** 6)]
    ref = O.prepare_test_pairs(n_boxes, max_pairs=10 ** 6)
    assert all(np.array_equal(a, b) for a, b in zip(got, ref))
    # cap + overlap on the largest supported filtered image (n = 128), ragged batch
    rng = np.random.default_rng(5)
    n_boxes = [128, 3, 40]
    boxes = [synth.make_boxes(rng, n, 800, 592) for n in n_boxes]
    scores = [rng.uniform(0.05, 1.0, n).astype(np.float32) for n in n_boxes]
    for overlap in (False, True):
        got = ops.enumerate_pairs(n_boxes, DEV, 2048, boxes=_t(np.concatenate(boxes)), scores=_t(np.concatenate(scores)),
                                  require_overlap=overlap)
        ref = O.prepare_test_pairs(n_boxes, 2048, scores=scores, boxes=boxes, require_overlap=overlap)
        assert all(np.array_equal(H.np_(a), b) for a, b in zip(got, ref)), overlap
    with pytest.raises(L.VetoError):
        ops.enumerate_pairs([200], DEV, 2048, scores=torch.ones(200, device=DEV))
    assert [tuple(p.shape) for p in ops.enumerate_pairs([], DEV)] == []


def test_globalize_pairs():
    n_boxes = [3, 1, 4]
    pairs = ops.enumerate_pairs(n_boxes, DEV)
    s, o = ops.globalize_pairs(pairs, n_boxes)
    rs, ro = O.global_pair_indices([H.np_(p) for p in pairs], n_boxes)
    assert np.array_equal(H.np_(s), rs) and np.array_equal(H.np_(o), ro)


# ------------------------------------------------------------------------------------------ a2/a3 gather
@pytest.mark.parametrize("name", ["cfg1_predcls_vg", "ragged_predcls", "sgdet_cap", "meet_gqa"])
def test_gather_bit_exact(name):
    c = CASES[name]
    batch = case_batch(c)
    g = load_golden(name)
    x2d, d2d, lv = ops.roi_gather([_t(f) for f in batch["feats"]], _t(batch["depth"]), _t(np.concatenate(batch["boxes"])),
                                  batch["n_boxes"], synth.POOLER_SCALES, synth.DEPTH_SCALE, return_levels=True)
    rx, rd = O.pooler_forward(batch["feats"], batch["depth"], batch["boxes"])
    x, d = H.np_(x2d), H.np_(d2d)
    assert np.array_equal(H.np_(lv), O.level_map(batch["boxes"]))
    assert np.array_equal(x, rx) and np.array_equal(d, rd)
    assert np.array_equal(x[:, ::16], g["x2d_sub"]) and np.array_equal(d[:, ::16], g["d2d_sub"])
    assert np.array_equal(x.sum(axis=(1, 2, 3), dtype=np.float64), g["x2d_sum"])


def test_roi_align_generic_and_edges():
    """The one-for-one _C.roi_align_forward replacement: boxes hanging outside the map, sub-pixel boxes, other
    pooled sizes / sampling ratios; and the scatter backward against the serial oracle."""
    rng = np.random.default_rng(9)
    B, C, Hh, Ww, scale = 2, 37, 20, 26, 0.125
    inp = rng.standard_normal((B, C, Hh, Ww), dtype=np.float32)
    W, Himg = Ww / scale, Hh / scale
    boxes = np.concatenate([synth.make_boxes(rng, 12, int(W), int(Himg), max_side=140.0),
                            np.array([[-30, -20, 40, 50], [W - 10, Himg - 10, W + 60, Himg + 40], [10, 10, 10.2, 10.1],
                                      [W + 50, Himg + 50, W + 90, Himg + 90]], np.float32)])
    rois = np.concatenate([rng.integers(0, B, (len(boxes), 1)).astype(np.float32), boxes], 1)
    for (ph, pw, sr) in [(8, 8, 2), (7, 7, 2), (4, 6, 3), (1, 1, 1)]:
        got = H.np_(ops.roi_align_forward(_t(inp), _t(rois), scale, ph, pw, sr))
        assert np.array_equal(got, O.roi_align(inp, rois, scale, ph, pw, sr)), (ph, pw, sr)
    grad = rng.standard_normal((len(rois), C, 8, 8), dtype=np.float32)
    gi = H.np_(ops.roi_align_backward(_t(grad), _t(rois), scale, 8, 8, B, C, Hh, Ww, 2))
    ref = O.roi_align_backward(grad, rois, (B, C, Hh, Ww), scale, 2)
    assert np.allclose(gi, ref, rtol=1e-4, atol=1e-5)   # fp32 atomics: summation order differs
    with pytest.raises(L.VetoError):
        ops.roi_align_forward(_t(inp), _t(rois), scale, 8, 8, 0)      # adaptive sampling is not on this path


# ------------------------------------------------------------------------------------------ a5-a10 head
def _golden_pairs(c, g, batch):
    """Pair lists in the golden row order (the reference's tie order over the cap)."""
    return [_t(p) for p in np.split(g["pairs"], np.cumsum(g["pair_counts"])[:-1])]


def _run_predictor(name, precision, chunk=0, pairs=None):
    c = CASES[name] if name in CASES else FULL_CASES[name]
    batch, state, g = case_batch(c), case_state(c), load_golden(name)
    cfg = H.make_cfg(c["predictor"], c["mode"], c["dataset"], c.get("max_pairs", 2048), c.get("require_overlap", False),
                     precision, chunk)
    cfg.ENSEMBLE_LEARNING.EXPERT_GROUP = bool(c.get("expert_group"))
    ds_obj = 151 if c["dataset"] == "VG" else 201
    bls = H.boxlists(batch, DEV, ds_obj)
    feats, depth = H.device_features(batch, DEV)
    from veto_b200 import registry
    fe = registry.make_roi_box_feature_extractor(cfg, 256, for_relation=True)
    pred = H.build_predictor(cfg, state, DEV)
    pairs = pairs if pairs is not None else _golden_pairs(c, g, batch)
    with torch.no_grad():
        x2d, d2d, a, b = fe(feats, bls, depth_features=depth)
        assert a is None and b is None
        out = pred(bls, pairs, None, None, roi_features=x2d, roi_depth_features=d2d)
    return c, batch, g, bls, pairs, pred, out


@pytest.mark.parametrize("precision", ["fp32", "bf16x3", "f16c8", "bf16", "f16"])
@pytest.mark.parametrize("name", ["cfg1_predcls_vg", "cfg1_default_init", "ragged_predcls", "sgdet_cap", "sgdet_overlap"])
def test_relation_logits(name, precision):
    c, batch, g, bls, pairs, pred, out = _run_predictor(name, precision)
    assert len(out) == 6                                           # relation_head.py:196 unpacks exactly six
    obj_dists, rel_dists, add_losses, incre, chosen, custom = out
    assert add_losses == {} and incre is None and chosen is None and custom is None
    assert [tuple(o.shape) for o in obj_dists] == [(n, 151) for n in batch["n_boxes"]]
    assert np.array_equal(np.concatenate([H.np_(o).argmax(1) for o in obj_dists]), g["obj_dists_argmax"])
    assert [r.shape[0] for r in rel_dists] == list(g["pair_counts"])
    logits, ref = np.concatenate([H.np_(r) for r in rel_dists]), g["logits"]
    assert rel_err(logits, ref) < TOL[precision]
    _check_argmax(logits, ref, precision)


def _check_argmax(logits, ref, precision):
    """argmax predicate labels: bit-exact in fp32 mode.  The tensor-core modes differ from the reference by their
    stated logit tolerance, so a label may only differ where the reference's own top-2 margin is inside that
    tolerance (a near-tie no reordering of fp32 sums is guaranteed to preserve either), and only rarely."""
    for sl in (slice(0, None), slice(1, None)):            # all classes, and the foreground classes PostProcessor uses
        a, b = logits[:, sl].argmax(1), ref[:, sl].argmax(1)
        if precision == "fp32":
            assert np.array_equal(a, b)
            continue
        bad = np.nonzero(a != b)[0]
        top2 = np.sort(ref[:, sl], 1)[:, -2:]
        margin = top2[:, 1] - top2[:, 0]
        assert np.all(margin[bad] <= TOL[precision] * np.abs(ref).max()), (precision, margin[bad])
        assert len(bad) <= (0.005 if precision in ("bf16x3", "f16c8") else 0.03) * len(a) + 1


@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
def test_tokens_match_reference(precision):
    """Encoder input [R,19,576] (factored patch embedding + gather/add) against tokens captured from the reference."""
    for name in ("cfg1_predcls_vg", "ragged_predcls"):
        c = CASES[name]
        batch, state, g = case_batch(c), case_state(c), load_golden(name)
        pred = H.build_predictor(H.make_cfg(precision=precision), state, DEV)
        x2d, d2d = ops.roi_gather([_t(f) for f in batch["feats"]], _t(batch["depth"]), _t(np.concatenate(batch["boxes"])),
                                  batch["n_boxes"], synth.POOLER_SCALES, synth.DEPTH_SCALE)
        pairs = _golden_pairs(c, g, batch)
        subj, obj = ops.globalize_pairs(pairs, batch["n_boxes"])
        pw = pred._pack(pred.rel_out.weight, pred.rel_out.bias)
        logits, feats, toks = ops.relation_forward(pw, _t(np.concatenate(batch["boxes"])), x2d, d2d, subj, obj,
                                                   labels=_t(np.concatenate(batch["labels"])), return_features=True,
                                                   return_tokens=True)
        tok = H.np_(toks)[g["token_rows"]]
        assert rel_err(tok, g["tokens"]) < 2e-5
        # rel_features is the CLS row the classifier consumed
        w, b = H.np_(pred.rel_out.weight).astype(np.float64), H.np_(pred.rel_out.bias).astype(np.float64)
        assert rel_err(H.np_(feats).astype(np.float64) @ w.T + b, H.np_(logits)) < 1e-5


@pytest.mark.parametrize("precision", ["fp32", "bf16x3", "f16c8", "bf16"])
@pytest.mark.parametrize("name", ["meet_gqa", "meet_vg", "meet_sgdet_nms", "meet_vg_experts"])
def test_meet_group_heads(name, precision):
    c, batch, g, bls, pairs, pred, out = _run_predictor(name, precision)
    obj_dists, rel_dists, add_losses, incre, chosen, custom = out
    assert isinstance(rel_dists, dict) and add_losses == {} and chosen is None and custom == {}
    assert list(incre) == list(g["incre_idx_list"])
    keys = [k[len("logits_"):] for k in g.files if k.startswith("logits_")]
    assert sorted(rel_dists) == sorted(keys)
    for k in keys:
        got, ref = H.np_(rel_dists[k]), g["logits_" + k]
        assert got.shape == ref.shape                           # un-split [R_total, n_k+2] (…:3843,3851-3853)
        assert rel_err(got, ref) < TOL[precision]
        _check_argmax(got, ref, precision)


@pytest.mark.parametrize("precision", ["fp32", "bf16x3", "f16c8"])
def test_meet_gqa_full_size_matches_reference(precision):
    """BASELINE.json configs[3] at its full size — VETOPredictor_MEET PredCls, GQA 201 / 101, 16 images x 20 boxes =
    6080 pairs — against the unmodified reference's outputs (tests/golden/meet_gqa_full.npz): pair indices and the ROI
    gather bit-exact, every 4th logit row of every group head within the mode's tolerance, and the argmax label of ALL
    6080 rows per head (a label may differ only where the reference's own top-2 margin is inside the tolerance)."""
    name = "meet_gqa_full"
    c = FULL_CASES[name]
    batch, g = case_batch(c), load_golden(name)
    enumerated = _pairs_gpu(c, batch)
    assert [len(p) for p in enumerated] == list(g["pair_counts"])
    assert np.array_equal(np.concatenate([H.np_(p) for p in enumerated]), g["pairs"])
    c, batch, g, bls, pairs, pred, out = _run_predictor(name, precision)
    obj_dists, rel_dists, add_losses, incre, chosen, custom = out
    assert list(incre) == list(g["incre_idx_list"]) and sum(len(p) for p in pairs) == 6080
    fe_cfg = H.make_cfg(c["predictor"], c["mode"], c["dataset"], precision=precision)
    from veto_b200 import registry
    feats, depth = H.device_features(batch, DEV)
    x2d, d2d, _, _ = registry.make_roi_box_feature_extractor(fe_cfg, 256, for_relation=True)(feats, bls, depth_features=depth)
    fs, rs = c["feat_stride"], c["row_stride"]
    assert np.array_equal(H.np_(x2d)[:, ::fs], g["x2d_sub"]) and np.array_equal(H.np_(d2d)[:, ::fs], g["d2d_sub"])
    scale = max(np.abs(g["logits_group_%d" % k]).max() for k in range(4))
    for k in range(4):
        got, ref = H.np_(rel_dists["group_%d" % k]), g["logits_group_%d" % k]
        assert got.shape[0] == 6080 and got[::rs].shape == ref.shape
        assert rel_err(got[::rs], ref) < TOL[precision]
        mine, theirs = got[:, 1:].argmax(1), g["argmax_group_%d" % k].astype(np.int64)
        bad = np.nonzero(mine != theirs)[0]
        if precision == "fp32":
            assert len(bad) == 0
        else:
            assert np.all(g["margin_group_%d" % k][bad] <= TOL[precision] * scale) and len(bad) <= 0.005 * 6080 + 1


def test_meet_per_class_nms_bit_exact():
    """veto_obj_nms_per_cls against Ensemble.nms_per_cls of the unmodified reference (golden: the score tile the
    reference fed it and the labels it returned) and against the numpy oracle on larger seeded cases, incl. ragged /
    empty images and a class count where many boxes compete."""
    from oracle import veto_oracle as O
    c, g = CASES["meet_sgdet_nms"], load_golden("meet_sgdet_nms")
    batch = case_batch(c, features=False)
    bpc = _t(np.concatenate(batch["boxes_per_cls"]))
    got = H.np_(ops.obj_nms_per_cls(_t(g["nms_scores"]), bpc, batch["n_boxes"], 0.5))
    assert np.array_equal(got, g["nms_labels"])
    assert (got != np.concatenate(batch["pred_labels"])).sum() >= 2
    # the drop-in's own score tile (torch softmax on the device) gives the same labels here
    onehot = torch.nn.functional.one_hot(_t(np.concatenate(batch["pred_labels"])), 151).float()
    assert np.array_equal(H.np_(ops.obj_nms_per_cls(torch.softmax(onehot, -1), bpc, batch["n_boxes"], 0.5)), g["nms_labels"])
    for seed, n_boxes, n_cls, thr in ((1, [80, 0, 33, 1], 4, 0.5), (2, [64] * 8, 2, 0.3), (3, [100, 7], 10, 0.7)):
        b = synth.make_batch(seed, n_boxes, H=592, W=800, mode="sgdet", features=False)
        synth.add_nms_fields(b, seed + 10, n_classes=n_cls, jitter=12.0)
        rng = np.random.default_rng(seed)
        # generic soft scores (not only one-hot): softmax of noisy logits peaked at the detector label
        lg = rng.standard_normal((sum(n_boxes), 151)).astype(np.float32)
        lg[np.arange(sum(n_boxes)), np.concatenate(b["pred_labels"])] += 4.0
        scores = O.softmax_rows(lg)
        ref = O.nms_per_cls(scores, b["boxes_per_cls"], b["n_boxes"], thr)
        got = H.np_(ops.obj_nms_per_cls(_t(scores), _t(np.concatenate(b["boxes_per_cls"])), b["n_boxes"], thr))
        assert np.array_equal(got, ref), (seed, int((got != ref).sum()))
        assert (ref != scores[:, 1:].argmax(1) + 1).any()          # suppression happened


@pytest.mark.parametrize("name", ["detsample_default", "detsample_tight"])
def test_detect_relsample(name):
    """RelationSampling.detect_relsample on the device (SGDet / SGCls training): binary matrices,
    locating_match and the row counts are identical to the unmodified reference's; the sampled rows satisfy the same
    structural properties against the oracle's candidate sets as the reference's own rows do (tests/test_oracle.py);
    reproducible under torch.manual_seed; the IoU-weighted draw prefers better-overlapping candidates."""
    from tests.cases import DETECT_SAMPLE_CASES
    from tests.train_util import check_detect_sample
    from veto_b200.sampling import make_roi_relation_samp_processor
    from veto_b200.structures import BoxList
    c, g = DETECT_SAMPLE_CASES[name], load_golden(name)
    imgs = synth.make_detect_case(c["seed"], c["n_tgt"])
    cfg = H.make_cfg(mode="sgdet")
    cfg.MODEL.ROI_RELATION_HEAD.BATCH_SIZE_PER_IMAGE = c["caps"][0]
    cfg.MODEL.ROI_RELATION_HEAD.POSITIVE_FRACTION = c["caps"][1]
    cfg.MODEL.ROI_RELATION_HEAD.REQUIRE_BOX_OVERLAP = c["require_overlap"]
    samp = make_roi_relation_samp_processor(cfg)
    assert samp.require_overlap == c["require_overlap"] and samp.num_sample_per_gt_rel == 4 and samp.fg_thres == 0.5
    batch, num_pos = c["caps"][0], int(c["caps"][0] * c["caps"][1])

    def run(seed):
        props, tgts = [], []
        for im in imgs:
            p = BoxList(_t(im["prp_boxes"]), im["size"], "xyxy")
            p.add_field("labels", _t(im["prp_labels"]))
            p.add_field("pred_scores", _t(im["prp_scores"]))
            t = BoxList(_t(im["tgt_boxes"]), im["size"], "xyxy")
            t.add_field("labels", _t(im["tgt_labels"]))
            t.add_field("relation", _t(im["relation"]))
            props.append(p)
            tgts.append(t)
        torch.manual_seed(seed)
        return samp.detect_relsample(props, tgts)

    props, rel_labels, rel_labels_all, rel_pairs, binarys = run(1)
    assert rel_labels_all is rel_labels
    cands = []
    for i, im in enumerate(imgs):
        cand = O.detect_relsample_candidates(im["prp_boxes"], im["prp_labels"], im["prp_scores"], im["tgt_boxes"],
                                             im["tgt_labels"], im["relation"], 0.5, c["require_overlap"])
        cands.append(cand)
        assert np.array_equal(H.np_(binarys[i]), g[f"binary/{i}"])
        assert np.array_equal(H.np_(props[i].get_field("locating_match")), g[f"locating_match/{i}"])
        pr, lb = H.np_(rel_pairs[i]), H.np_(rel_labels[i])
        assert pr.dtype == np.int64 and lb.dtype == np.int64
        check_detect_sample(cand, pr, lb, batch, num_pos)
        assert len(pr) == len(g[f"pairs/{i}"]) and int((lb > 0).sum()) == int((g[f"labels/{i}"] > 0).sum())
    again, other = run(1), run(2)
    assert all(torch.equal(a, b) for a, b in zip(again[3], rel_pairs))
    assert any(not torch.equal(a, b) for a, b in zip(other[3], rel_pairs))
    if name == "detsample_default":
        # the weighted draw: over many seeds, among the candidates of a ground-truth relation with more than four of
        # them, the selection frequency follows the weight iou_head * iou_tail (rank correlation), and every draw keeps 4
        i, rel = next((i, r) for i, cnd in enumerate(cands) for r in cnd["gt"] if len(r[3]) > 6)
        h, t, l, cc = rel
        w = np.array([cands[i]["ious"][h, a] * cands[i]["ious"][t, b] for a, b in cc])
        freq = np.zeros(len(cc))
        trials = 200
        for s in range(trials):
            out = run(100 + s)
            pr, lb = H.np_(out[3][i]), H.np_(out[1][i])
            rows = {(int(a), int(b)) for (a, b), x in zip(pr, lb) if x == l}
            hit = np.array([(a, b) in rows for a, b in cc], float)
            freq += hit
        others = sum(1 for r in cands[i]["gt"] if r[2] == l and r is not rel)
        if others == 0:
            assert np.all(np.isclose(freq.sum(), 4 * trials))
        order_w, order_f = np.argsort(np.argsort(w)), np.argsort(np.argsort(freq))
        rho = np.corrcoef(order_w, order_f)[0, 1]
        assert rho > 0.5, (rho, w, freq)


def test_recall_evaluation_matches_reference():
    """veto_sgg_match + evaluation.recall_at_k against SGRecall.calculate_recall of the unmodified
    reference: per ground-truth triplet the rank of the first matching prediction, per prediction the number of
    matches, recall@20/50/100 per image and the per-predicate hit counts — all exact."""
    from tests.cases import EVAL_CASES
    from veto_b200 import evaluation as E
    from veto_b200.structures import BoxList
    c, g = EVAL_CASES["eval_recall"], load_golden("eval_recall")
    imgs = synth.make_eval_case(c["seed"], c["n_objs"], c["n_gt_rels"], c["n_pred_rels"])
    preds, gts = [], []
    for im in imgs:
        gt = BoxList(_t(im["boxes"]), im["size"], "xyxy")
        gt.add_field("labels", _t(im["labels"]))
        gt.add_field("relation_tuple", _t(im["relation_tuple"]))
        pr = BoxList(_t(im["pred_boxes"]), im["size"], "xyxy")
        pr.add_field("pred_labels", _t(im["pred_labels"]))
        pr.add_field("rel_pair_idxs", _t(im["rel_pair_idxs"]))
        pr.add_field("pred_rel_scores", _t(im["pred_rel_scores"]))
        pr.add_field("pred_scores", _t(im["pred_scores"]))
        preds.append(pr)
        gts.append(gt)
    # an image without ground-truth relations is skipped (vg_eval.py:473-474)
    empty = BoxList(_t(imgs[0]["boxes"]), imgs[0]["size"], "xyxy")
    empty.add_field("labels", _t(imgs[0]["labels"]))
    empty.add_field("relation_tuple", torch.zeros((0, 3), dtype=torch.int64, device=DEV))
    out = E.recall_at_k(preds[:2] + [preds[0]] + preds[2:], gts[:2] + [empty] + gts[2:])
    assert len(out["first_match"]) == len(imgs)
    for i in range(len(imgs)):
        assert np.array_equal(out["first_match"][i].numpy().astype(np.int64), g[f"first_match/{i}"])
    for k in (20, 50, 100):
        assert np.array_equal(np.array(out["recall"][k]), g[f"recall/{k}"])
        per = np.array([[r, h, n] for r, (h, n) in sorted(out["hits_per_rel"][k].items())], np.int64)
        assert np.array_equal(per, g[f"per_rel/{k}"])
    ng = E.recall_nogc_at_k(preds, gts)                                             # SGNoGraphConstraintRecall (:213-252)
    for i in range(len(imgs)):
        assert np.array_equal(ng["first_match"][i].numpy().astype(np.int64), g[f"nogc_first_match/{i}"])
    for k in (20, 50, 100):
        assert np.array_equal(np.array(ng["recall"][k]), g[f"recall_nogc/{k}"])
    assert max(ng["recall"][100]) > max(out["recall"][100]) - 1e-9                  # no graph constraint can only help
    ngm = E.mean_recall(ng["first_match"], out["gt_predicates"], 51)                # SGNGMeanRecall (:470-548)
    for k in (20, 50, 100):
        assert abs(ngm["mean_recall"][k] - float(g[f"ng_mean_recall/{k}"])) <= 1e-12
    zs = E.zeroshot_recall(out["first_match"], gts[:2] + [empty] + gts[2:], _t(g["zeroshot_triplets"]))   # SGZeroShotRecall
    pa = E.pair_accuracy(preds, gts, predcls_like=False)                            # SGPairAccuracy (:338-366)
    for k in (20, 50, 100):
        assert np.array_equal(np.array(zs[k]), g[f"zeroshot_recall/{k}"])
        assert np.array_equal(np.array(pa["hit"][k]), g[f"accuracy_hit/{k}"])
        assert np.array_equal(np.array(pa["count"][k]), g[f"accuracy_count/{k}"])
    assert sum(pa["hit"][100]) > 0 and len(zs[100]) == len(imgs)
    mr = E.mean_recall(out["first_match"], out["gt_predicates"], 51)                # SGMeanRecall (:424-466)
    for k in (20, 50, 100):
        assert np.allclose(mr["mean_recall_list"][k], g[f"mean_recall_list/{k}"], rtol=1e-12, atol=0)
        assert abs(mr["mean_recall"][k] - float(g[f"mean_recall/{k}"])) <= 1e-12
    # raw kernel outputs incl. the per-prediction match counts
    gt_t, gt_b, pr_t, pr_b = [], [], [], []
    for pr, gt in zip(preds, gts):
        rt = gt.get_field("relation_tuple")
        t, b = E.triplets(rt[:, :2], rt[:, 2], gt.get_field("labels"), gt.bbox)
        gt_t.append(t); gt_b.append(b)
        lab = 1 + pr.get_field("pred_rel_scores")[:, 1:].argmax(1)
        t, b = E.triplets(pr.get_field("rel_pair_idxs"), lab, pr.get_field("pred_labels"), pr.bbox)
        pr_t.append(t); pr_b.append(b)
    first, hits = ops.sgg_match(torch.cat(gt_t), torch.cat(gt_b), [len(t) for t in gt_t], torch.cat(pr_t), torch.cat(pr_b),
                                [len(t) for t in pr_t], 0.5)
    assert np.array_equal(H.np_(hits), np.concatenate([g[f"pred_hits/{i}"] for i in range(len(imgs))]))
    assert np.array_equal(H.np_(first).astype(np.int64), np.concatenate([g[f"first_match/{i}"] for i in range(len(imgs))]))
    # a larger seeded batch against the numpy oracle
    big = synth.make_eval_case(77, [40, 33, 25, 60, 12, 48], n_gt_rels=40, n_pred_rels=1000)
    gt_t, gt_b, pr_t, pr_b, ref_first = [], [], [], [], []
    for im in big:
        s, o, p = im["relation_tuple"][:, 0], im["relation_tuple"][:, 1], im["relation_tuple"][:, 2]
        gtt, gtb = np.column_stack((im["labels"][s], p, im["labels"][o])), np.column_stack((im["boxes"][s], im["boxes"][o]))
        ps, po = im["rel_pair_idxs"][:, 0], im["rel_pair_idxs"][:, 1]
        prt = np.column_stack((im["pred_labels"][ps], 1 + im["pred_rel_scores"][:, 1:].argmax(1), im["pred_labels"][po]))
        prb = np.column_stack((im["pred_boxes"][ps], im["pred_boxes"][po]))
        p2g = O.compute_pred_matches(gtt, prt, gtb, prb, 0.5)
        f = np.full(len(gtt), 2 ** 31 - 1, np.int64)
        for k, gs in enumerate(p2g):
            for gi in gs:
                f[gi] = min(f[gi], k)
        ref_first.append(f)
        gt_t.append(gtt); gt_b.append(gtb); pr_t.append(prt); pr_b.append(prb)
    first, _ = ops.sgg_match(_t(np.concatenate(gt_t)), _t(np.concatenate(gt_b)), [len(t) for t in gt_t],
                             _t(np.concatenate(pr_t)), _t(np.concatenate(pr_b)), [len(t) for t in pr_t], 0.5)
    assert np.array_equal(H.np_(first).astype(np.int64), np.concatenate(ref_first))
    assert (np.concatenate(ref_first) < 2 ** 31 - 1).mean() > 0.3


@pytest.mark.parametrize("name", ["relsample_under_caps", "relsample_over_caps"])
def test_gtbox_relsample(name):
    """RelationSampling.gtbox_relsample on the device against the unmodified reference's output on
    the same relation matrices and against the oracle's candidate sets.  The random draws come from a different
    stream than torch.randperm, so parity is: identical foreground rows / labels and an identical background SET
    when everything fits; identical row counts, subset-of-candidates, no duplicates and consistent labels when the
    caps bite; plus reproducibility under torch.manual_seed and a uniformity check of the random subset."""
    from tests.cases import RELSAMPLE_CASES
    from veto_b200.sampling import make_roi_relation_samp_processor
    from veto_b200.structures import BoxList
    c, g = RELSAMPLE_CASES[name], load_golden(name)
    mats = synth.make_relation_matrices(c["seed"], c["n_boxes"], 51, c["fg_per_image"])
    cfg = H.make_cfg()
    cfg.MODEL.ROI_RELATION_HEAD.BATCH_SIZE_PER_IMAGE = c["caps"][0]
    cfg.MODEL.ROI_RELATION_HEAD.POSITIVE_FRACTION = c["caps"][1]
    samp = make_roi_relation_samp_processor(cfg)
    batch, num_pos = c["caps"][0], int(c["caps"][0] * c["caps"][1])

    def run(seed):
        props, tgts = [], []
        for n, m in zip(c["n_boxes"], mats):
            box = torch.rand(n, 4, device=DEV) * 100
            props.append(BoxList(box, (416, 320), "xyxy"))
            t = BoxList(box.clone(), (416, 320), "xyxy")
            t.add_field("relation", _t(m))
            tgts.append(t)
        torch.manual_seed(seed)
        return samp.gtbox_relsample(props, tgts)

    props, rel_labels, rel_pairs, binarys = run(1)
    for i, m in enumerate(mats):
        fg, labels, bg, binary = O.gtbox_relsample_candidates(m)
        pr, lb = H.np_(rel_pairs[i]).reshape(-1, 2), H.np_(rel_labels[i])
        ref_pairs, ref_labels = g[f"pairs/{i}"].reshape(-1, 2), g[f"labels/{i}"]
        assert pr.dtype == np.int64 and lb.dtype == np.int64
        assert np.array_equal(H.np_(binarys[i]), g[f"binary/{i}"])
        assert np.array_equal(H.np_(props[i].get_field("locating_match")), g[f"locating_match/{i}"])
        assert len(pr) == len(ref_pairs) and int((lb > 0).sum()) == int((ref_labels > 0).sum())
        n_fg = int((ref_labels > 0).sum())
        cand = {tuple(p): int(l) for p, l in zip(fg, labels)}
        if len(fg) <= num_pos:
            assert np.array_equal(pr[:n_fg], ref_pairs[:n_fg]) and np.array_equal(lb[:n_fg], ref_labels[:n_fg])
            assert {tuple(p) for p in pr[n_fg:]} == {tuple(p) for p in ref_pairs[n_fg:]} or len(bg) > batch - n_fg
        assert all(cand[tuple(p)] == int(l) for p, l in zip(pr[:n_fg], lb[:n_fg]))
        assert len({tuple(p) for p in pr}) == len(pr)                                  # no duplicates
        assert {tuple(p) for p in pr[n_fg:]} <= {tuple(p) for p in bg} and np.all(lb[n_fg:] == 0)
    # reproducible under the same torch seed, different under another
    again = run(1)
    other = run(2)
    assert all(torch.equal(a, b) for a, b in zip(again[2], rel_pairs))
    assert any(not torch.equal(a, b) for a, b in zip(other[2], rel_pairs) if len(a) > 2)
    if name == "relsample_over_caps":
        # every background candidate of image 0 is kept about equally often: 24 of 366 slots -> p = 0.0656
        fg, labels, bg, _ = O.gtbox_relsample_candidates(mats[0])
        hits = {tuple(p): 0 for p in bg}
        trials = 300
        for s in range(trials):
            out = run(100 + s)
            for p in H.np_(out[2][0])[num_pos:]:
                hits[tuple(p)] += 1
        kept = batch - num_pos
        p_exp = kept / len(bg)
        freq = np.array(list(hits.values())) / trials
        assert abs(freq.mean() - p_exp) < 1e-9
        sigma = np.sqrt(p_exp * (1 - p_exp) / trials)
        assert freq.max() < p_exp + 6 * sigma and freq.min() > p_exp - 6 * sigma, (freq.min(), freq.max(), p_exp)


@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
def test_meet_postprocess_ensemble_merge(precision):
    """PostProcessor MEET 'ensemble' branch (inference.py:284-397): the merged, ranked candidate list of the unmodified
    reference for a one-image batch (the only thing it supports), and the batch generalisation against the oracle."""
    from veto_b200.postprocess import make_roi_relation_post_processor
    name = "meet_vg"
    c, batch, g, bls, pairs, pred, out = _run_predictor(name, precision)
    cfg = H.make_cfg("VETOPredictor_MEET", "predcls", "VG", precision=precision)
    post = make_roi_relation_post_processor(cfg)
    res = post((out[1], [b.get_field("predict_logits") for b in bls]), pairs, bls, incre_idx_list=out[3])
    assert len(res) == 1
    pp, probs = H.np_(res[0].get_field("rel_pair_idxs")), H.np_(res[0].get_field("pred_rel_scores"))
    labels, trip = H.np_(res[0].get_field("pred_rel_labels")), H.np_(res[0].get_field("triple_scores")).astype(np.float64)
    assert pp.dtype == np.float32 and pp.shape == g["mpost_pairs"].shape and probs.shape == g["mpost_probs"].shape
    assert np.all(np.diff(trip) <= 0)
    gap = np.full(len(trip), np.inf)
    gap[1:] = np.minimum(gap[1:], trip[:-1] - trip[1:])
    gap[:-1] = np.minimum(gap[:-1], trip[:-1] - trip[1:])
    # rows whose score is separated from both neighbours by more than the mode's logit error must sit at the same rank
    clear = gap > {"fp32": 2e-5, "bf16x3": 4e-4}[precision] * trip
    assert clear.mean() > 0.3, clear.mean()
    assert np.array_equal(pp[clear], g["mpost_pairs"][clear]) and np.array_equal(labels[clear], g["mpost_labels"][clear])
    assert np.abs(probs[clear] - g["mpost_probs"][clear]).max() <= TOL[precision]
    assert np.array_equal(probs[clear] != 0, g["mpost_probs"][clear] != 0)          # global-column scatter pattern
    # evaluation re-derives the global label from the scattered probabilities (sgg_eval.py:150): same as the reference's
    assert np.array_equal(probs[clear][:, 1:].argmax(1), g["mpost_probs"][clear][:, 1:].argmax(1))
    # batch of two images: per image identical to the oracle's single-image merge of OUR logits
    name = "meet_gqa"
    c, batch, g, bls, pairs, pred, out = _run_predictor(name, precision)
    cfg = H.make_cfg("VETOPredictor_MEET", "predcls", "GQA", precision=precision)
    res = make_roi_relation_post_processor(cfg)((out[1], [b.get_field("predict_logits") for b in bls]), pairs, bls,
                                                incre_idx_list=out[3])
    off = 0
    for i, r in enumerate(res):
        n = len(pairs[i])
        gl = {k: H.np_(v)[off:off + n] for k, v in out[1].items()}
        off += n
        ora = O.postprocess_meet(gl, H.np_(bls[i].get_field("predict_logits")), H.np_(pairs[i]), list(out[3]))
        s = ora["triple"].astype(np.float64)
        gap = np.full(len(s), np.inf)
        gap[1:] = np.minimum(gap[1:], s[:-1] - s[1:])
        gap[:-1] = np.minimum(gap[:-1], s[:-1] - s[1:])
        clear = gap > 1e-5 * s
        assert len(s) == 4 * n and clear.mean() > 0.5
        assert np.array_equal(H.np_(r.get_field("rel_pair_idxs"))[clear].astype(np.int64), ora["pairs"][clear])
        assert np.array_equal(H.np_(r.get_field("pred_rel_labels"))[clear], ora["labels"][clear])
        assert np.abs(H.np_(r.get_field("pred_rel_scores"))[clear] - ora["probs"][clear]).max() <= 1e-5


@pytest.mark.parametrize("voting", ["C", "U"])
@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
def test_meet_postprocess_expert_voting(precision, voting):
    """PostProcessor EXPERT_GROUP branch (inference.py:93-283): three experts per group vote on every candidate —
    survivors, their head-local labels, scattered probabilities and ranking against the unmodified reference; and, on
    OUR logits, exactly the oracle's survivors."""
    from veto_b200.postprocess import make_roi_relation_post_processor
    name = "meet_vg_experts"
    c, batch, g, bls, pairs, pred, out = _run_predictor(name, precision)
    assert sorted(out[1]) == sorted("group_%d%d" % (k, e) for k in range(5) for e in (1, 2, 3))
    cfg = H.make_cfg("VETOPredictor_MEET", "predcls", "VG", precision=precision)
    cfg.ENSEMBLE_LEARNING.EXPERT_GROUP = True
    cfg.ENSEMBLE_LEARNING.VOTING = voting
    res = make_roi_relation_post_processor(cfg)((out[1], [b.get_field("predict_logits") for b in bls]), pairs, bls,
                                                incre_idx_list=out[3])
    pp, probs = H.np_(res[0].get_field("rel_pair_idxs")), H.np_(res[0].get_field("pred_rel_scores"))
    labels, trip = H.np_(res[0].get_field("pred_rel_labels")), H.np_(res[0].get_field("triple_scores")).astype(np.float64)
    assert pp.dtype == np.float32 and np.all(np.diff(trip) <= 0)
    # against the oracle's vote on OUR logits: the same survivors in the same order wherever scores are distinct
    ora = O.postprocess_meet_vote({k: H.np_(v) for k, v in out[1].items()}, H.np_(bls[0].get_field("predict_logits")),
                                  H.np_(pairs[0]), list(out[3]), voting)
    assert len(trip) == len(ora["triple"])
    so = ora["triple"].astype(np.float64)
    gap = np.full(len(so), np.inf)
    gap[1:] = np.minimum(gap[1:], so[:-1] - so[1:])
    gap[:-1] = np.minimum(gap[:-1], so[:-1] - so[1:])
    clear = gap > 1e-5 * so
    assert clear.mean() > 0.5
    assert np.array_equal(pp[clear].astype(np.int64), ora["pairs"][clear]) and np.array_equal(labels[clear], ora["labels"][clear])
    assert np.abs(probs[clear] - ora["probs"][clear]).max() <= 1e-5
    # against the reference's output: the survivor count may differ only by candidates whose experts' top-2 margin is
    # inside the mode's logit tolerance; rows with clearly separated scores sit at the same rank with the same content
    ref_pairs, ref_probs, ref_labels = g[f"vote{voting}_pairs"], g[f"vote{voting}_probs"], g[f"vote{voting}_labels"]
    assert abs(len(trip) - len(ref_labels)) <= (0 if precision == "fp32" else 3)
    if len(trip) == len(ref_labels):
        gap = np.full(len(trip), np.inf)
        gap[1:] = np.minimum(gap[1:], trip[:-1] - trip[1:])
        gap[:-1] = np.minimum(gap[:-1], trip[:-1] - trip[1:])
        clear = gap > {"fp32": 2e-5, "bf16x3": 4e-4}[precision] * trip
        assert clear.mean() > 0.3
        assert np.array_equal(pp[clear], ref_pairs[clear]) and np.array_equal(labels[clear], ref_labels[clear])
        assert np.abs(probs[clear] - ref_probs[clear]).max() <= TOL[precision]


@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
def test_sgdet_postprocess_late_nms(precision):
    """PostProcessor at SGDet test time (inference.py:398-453 with use_gt_box False): object labels from the late
    per-class NMS (obj_prediction_nms), scores, per-class regressed boxes and the triple ranking against the
    unmodified reference; the NMS kernel bit-exact on the reference's own softmax tile and on larger seeded cases."""
    name = "sgdet_post_nms"
    c, batch, g, bls, pairs, pred, out = _run_predictor(name, precision)
    bpc = _t(np.concatenate(batch["boxes_per_cls"]))
    prob = g["post_obj_prob"].copy()
    got = H.np_(ops.obj_nms_per_cls(_t(prob), bpc, batch["n_boxes"], 0.5, late_nms=True))
    assert np.array_equal(got, g["post_obj_labels"])
    from veto_b200.postprocess import make_roi_relation_post_processor
    cfg = H.make_cfg("VETOPredictor", "sgdet", precision=precision)
    cfg.TEST.RELATION.LATER_NMS_PREDICTION_THRES = 0.5
    res = make_roi_relation_post_processor(cfg)((out[1], [b.get_field("predict_logits") for b in bls]), pairs, bls)
    assert np.array_equal(np.concatenate([H.np_(r.get_field("pred_labels")) for r in res]), g["post_obj_labels"])
    assert np.allclose(np.concatenate([H.np_(r.get_field("pred_scores")) for r in res]), g["post_obj_scores"], rtol=1e-5)
    assert np.array_equal(np.concatenate([H.np_(r.bbox) for r in res]), g["post_boxes"])
    assert all(not r.has_field("boxes_per_cls") for r in res)          # a NEW BoxList (inference.py:431)
    pp = np.concatenate([H.np_(r.get_field("rel_pair_idxs")) for r in res])
    labels = np.concatenate([H.np_(r.get_field("pred_rel_labels")) for r in res])
    trip = np.concatenate([H.np_(r.get_field("triple_scores")) for r in res]).astype(np.float64)
    # rows whose triple score is separated from its neighbours by more than the logit tolerance must agree
    gap = np.full(len(trip), np.inf)
    off = 0
    for n in g["pair_counts"]:
        t = trip[off:off + n]
        assert np.all(np.diff(t) <= 0)
        gp = np.full(n, np.inf)
        gp[1:] = np.minimum(gp[1:], t[:-1] - t[1:])
        gp[:-1] = np.minimum(gp[:-1], t[:-1] - t[1:])
        gap[off:off + n] = gp
        off += n
    clear = gap > 4e-3 * trip
    assert clear.mean() > 0.5
    assert np.array_equal(pp[clear], g["post_pairs"][clear]) and np.array_equal(labels[clear], g["post_labels"][clear])
    # larger seeded cases against the numpy oracle, generic soft scores
    for seed, n_boxes, thr in ((5, [80, 0, 33, 1], 0.5), (6, [64] * 4, 0.3)):
        b = synth.make_batch(seed, n_boxes, H=592, W=800, mode="sgdet", features=False)
        synth.add_nms_fields(b, seed + 10, n_classes=4, jitter=12.0, peak=2.0)
        ref, got_all, o = [], [], 0
        scores = np.concatenate([O.softmax_rows(l) for l in b["predict_logits"] if len(l)])
        for l, bx in zip(b["predict_logits"], b["boxes_per_cls"]):
            if len(l):
                ref.append(O.obj_prediction_nms(O.softmax_rows(l), bx, thr))
        got = H.np_(ops.obj_nms_per_cls(_t(scores), _t(np.concatenate(b["boxes_per_cls"])), b["n_boxes"], thr, late_nms=True))
        assert np.array_equal(got, np.concatenate(ref)), seed


@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
def test_top50_ranking_and_postprocess(precision):
    """PostProcessor parity: triple scores, and the top-50 triplet ranking of the reference (tie-aware)."""
    name = "cfg1_predcls_vg"
    c, batch, g, bls, pairs, pred, out = _run_predictor(name, precision)
    from veto_b200.postprocess import make_roi_relation_post_processor
    cfg = H.make_cfg(precision=precision)
    res = make_roi_relation_post_processor(cfg)((out[1], [b.get_field("predict_logits") for b in bls]), pairs, bls)
    pp = np.concatenate([H.np_(r.get_field("rel_pair_idxs")) for r in res])
    labels = np.concatenate([H.np_(r.get_field("pred_rel_labels")) for r in res])
    probs = np.concatenate([H.np_(r.get_field("pred_rel_scores")) for r in res])
    scores = probs[:, 1:].max(1)
    assert pp.dtype == np.int64 and labels.dtype == np.int64
    assert np.allclose(scores, g["post_scores"], rtol=5e-4)
    assert np.all(np.diff(np.concatenate([H.np_(r.get_field("triple_scores")) for r in res])) <= 0)
    assert np.allclose(probs.sum(1), 1.0, atol=1e-5)
    # rows whose reference score is separated from its neighbours by more than the logit tolerance must agree
    s = g["post_scores"].astype(np.float64)
    gap = np.full(len(s), np.inf)
    gap[1:] = np.minimum(gap[1:], s[:-1] - s[1:])
    gap[:-1] = np.minimum(gap[:-1], s[:-1] - s[1:])
    clear = gap > 2e-3 * s
    assert np.array_equal(pp[clear], g["post_pairs"][clear]) and np.array_equal(labels[clear], g["post_labels"][clear])
    # top-50 as a set of (subject, object, label) triplets, allowing swaps only among unclear rows
    top_ref = {tuple(r) + (int(l),) for r, l in zip(g["post_pairs"][:50], g["post_labels"][:50])}
    top_got = {tuple(r) + (int(l),) for r, l in zip(pp[:50], labels[:50])}
    assert len(top_ref - top_got) <= int((~clear[45:55]).sum())
    # and against the oracle's post-processing of OUR logits: exact ranking wherever scores are distinct
    logits = np.concatenate([H.np_(r) for r in out[1]])
    obj_logits = [H.np_(b.get_field("predict_logits")) for b in bls]
    ora = O.postprocess(np.split(logits, np.cumsum(g["pair_counts"])[:-1]), obj_logits, [H.np_(p) for p in pairs])
    so = np.concatenate([r["triple_scores"] for r in ora]).astype(np.float64)
    gap = np.full(len(so), np.inf)
    gap[1:] = np.minimum(gap[1:], so[:-1] - so[1:])
    gap[:-1] = np.minimum(gap[:-1], so[:-1] - so[1:])
    clear = gap > 1e-5 * so
    assert np.array_equal(pp[clear], np.concatenate([r["rel_pair_idxs"] for r in ora])[clear])


def _clear_ranks(sorted_scores, rel_gap=1e-5):
    """Rows of a descending score list whose score is separated from both neighbours by more than rel_gap (expf of the
    device vs numpy's exp differ in the last bits): only these must sit at the same rank on both sides."""
    s = np.asarray(sorted_scores, dtype=np.float64)
    gap = np.full(len(s), np.inf)
    gap[1:] = np.minimum(gap[1:], s[:-1] - s[1:])
    gap[:-1] = np.minimum(gap[:-1], s[:-1] - s[1:])
    return gap > rel_gap * s


def test_postprocess_oversized_images():
    """Images whose ranked list does not fit the shared-memory sort (more than 16 384 rows: MAX_PROPOSAL_PAIR 4096+ with
    the MEET heads merged, or a vanilla image with more than 16 384 pairs) sort in global memory: the same
    results as the oracle, next to a small image in the same batch."""
    rng = np.random.default_rng(3)
    n_boxes = [150, 12]                                         # 150 * 149 = 22 350 pairs > 16 384
    pairs = O.prepare_test_pairs(n_boxes, max_pairs=1 << 30)
    counts = [len(p) for p in pairs]
    R = sum(counts)
    logits = rng.standard_normal((R, 51)).astype(np.float32) * 3
    obj_logits = [rng.standard_normal((n, 151)).astype(np.float32) * 3 for n in n_boxes]
    ora = O.postprocess(np.split(logits, np.cumsum(counts)[:-1]), obj_logits, pairs)
    obj_scores = np.concatenate([o["pred_scores"] for o in ora])
    po, pr, lab, tri = ops.postprocess(_t(logits), _t(np.concatenate(pairs)), _t(obj_scores), counts, n_boxes)
    off = 0
    for o, n in zip(ora, counts):
        t = H.np_(tri)[off:off + n]
        assert np.all(np.diff(t) <= 0) and np.allclose(t, o["triple_scores"], rtol=2e-6)
        keep = _clear_ranks(o["triple_scores"])
        assert keep.mean() > 0.5
        assert np.array_equal(H.np_(po)[off:off + n][keep], o["rel_pair_idxs"][keep])
        assert np.array_equal(H.np_(lab)[off:off + n][keep], o["pred_rel_labels"][keep])
        assert np.abs(H.np_(pr)[off:off + n][keep] - o["pred_rel_scores"][keep]).max() < 1e-6
        off += n
    # MEET merge: 5 heads x 3540 pairs = 17 700 merged rows
    sizes = synth.GROUP_SPLITS[("VG", "divide4")]
    from veto_b200.predictor import incre_idx_list
    incre = incre_idx_list(sizes, 51)
    n_boxes = [60]
    pairs = O.prepare_test_pairs(n_boxes, max_pairs=1 << 30)[0]
    heads = [n + 2 for n in sizes]
    gl = {"group_%d" % k: rng.standard_normal((len(pairs), n)).astype(np.float32) * 2 for k, n in enumerate(heads)}
    obj_logit = rng.standard_normal((60, 151)).astype(np.float32) * 3
    ora = O.postprocess_meet(gl, obj_logit, pairs, incre)
    op = O.softmax_rows(obj_logit)
    op[:, 0] = 0
    col_map = []
    for k, n in enumerate(heads):
        col_map += [0] + [c for c, g in enumerate(incre) if g == k + 1] + [0]
    po, pr, lab, tri = ops.postprocess_meet(_t(np.concatenate([gl["group_%d" % k] for k in range(5)], 1)), heads, col_map, 51,
                                            _t(pairs), _t(op[:, 1:].max(1)), [len(pairs)], n_boxes)
    t, s = H.np_(tri), ora["triple"]
    assert len(t) == 5 * len(pairs) == 17700 and np.all(np.diff(t) <= 0) and np.allclose(t, s, rtol=2e-6)
    keep = _clear_ranks(s)
    assert keep.mean() > 0.5
    assert np.array_equal(H.np_(po)[keep], ora["pairs"][keep]) and np.array_equal(H.np_(lab)[keep], ora["labels"][keep])
    assert np.abs(H.np_(pr)[keep] - ora["probs"][keep]).max() < 1e-6


def test_chunking_and_permutation_invariance():
    """Size-independent properties: the chunk size never changes a bit of the result, and permuting the pair list
    permutes the logits (rows are independent)."""
    name = "cfg1_predcls_vg"
    for precision in ("fp32", "bf16x3"):
        base = np.concatenate([H.np_(r) for r in _run_predictor(name, precision)[6][1]])
        for chunk in (1, 37, 380):
            alt = np.concatenate([H.np_(r) for r in _run_predictor(name, precision, chunk=chunk)[6][1]])
            assert np.array_equal(base, alt), (precision, chunk)
        c = CASES[name]
        g = load_golden(name)
        perm = np.random.default_rng(3).permutation(len(g["pairs"]))
        out = _run_predictor(name, precision, pairs=[_t(g["pairs"][perm])])[6]
        assert np.array_equal(H.np_(out[1][0]), base[perm]), precision


def test_frequency_bias_epilogue_off_by_default_and_additive():
    """The optional freq-bias add (model_motifs.py:29-38 semantics) is OFF for VETO parity; when enabled it adds
    table[label_s * num_obj + label_o]."""
    name = "ragged_predcls"
    c, batch, g, bls, pairs, pred, out = _run_predictor(name, "fp32")
    base = torch.cat(list(out[1]))
    assert pred.use_freq_bias is False
    table = torch.randn(151 * 151, 51, device=DEV)
    pred.use_freq_bias, pred.freq_bias_table = True, table
    feats, depth = H.device_features(batch, DEV)
    x2d, d2d = ops.roi_gather(feats[:4], depth, torch.cat([b.bbox for b in bls]), batch["n_boxes"], synth.POOLER_SCALES,
                              synth.DEPTH_SCALE)
    with torch.no_grad():
        biased = torch.cat(list(pred(bls, pairs, None, None, roi_features=x2d, roi_depth_features=d2d)[1]))
    lab = torch.cat([b.get_field("labels") for b in bls])
    s, o = ops.globalize_pairs(pairs, batch["n_boxes"])
    assert torch.equal(biased, base + table[lab[s.long()] * 151 + lab[o.long()]])


def test_full_size_sgdet_batch_properties():
    """BASELINE.json configs[2] shape (80 proposals / image, 6320 pairs / image, cap 8192) on a few images: pair
    enumeration against the closed form, finite logits, mirror/permutation consistency, bf16x3 vs fp32 agreement."""
    n_img, n_box = 3, 80
    batch = synth.make_batch(31, [n_box] * n_img, mode="sgdet")
    state = synth.predictor_state(13)
    outs = {}
    for precision in ("fp32", "bf16x3"):
        cfg = H.make_cfg(mode="sgdet", max_pairs=8192, precision=precision)
        outs[precision] = H.run_head(cfg, state, batch, DEV, post=False)
    pairs = [H.np_(p) for p in outs["fp32"]["pairs"]]
    assert [len(p) for p in pairs] == [n_box * (n_box - 1)] * n_img
    r = np.arange(n_box * (n_box - 1))
    i, j = r // (n_box - 1), r % (n_box - 1)
    assert all(np.array_equal(p, np.stack([i, j + (j >= i)], 1)) for p in pairs)
    a = np.concatenate([H.np_(x) for x in outs["fp32"]["rel_dists"]])
    b = np.concatenate([H.np_(x) for x in outs["bf16x3"]["rel_dists"]])
    assert np.isfinite(a).all() and a.shape == (n_img * n_box * (n_box - 1), 51)
    assert rel_err(b, a) < 2e-4
    assert (a.argmax(1) == b.argmax(1)).mean() > 0.999
    # oracle spot-check of a random sample of pairs at full batch size
    sel = np.random.default_rng(0).choice(len(a), 48, replace=False)
    x_ref, d_ref = O.pooler_forward(batch["feats"], batch["depth"], batch["boxes"])
    assert np.array_equal(H.np_(outs["fp32"]["x2d"]), x_ref)
    off = np.repeat(np.arange(n_img) * n_box, n_box * (n_box - 1))
    gp = np.concatenate(pairs) + off[:, None]
    one = dict(batch, n_boxes=[n_img * n_box], boxes=[np.concatenate(batch["boxes"])],
               predict_logits=[np.concatenate(batch["predict_logits"])])
    ref = O.predictor_forward(state, one, [gp[sel]], x_ref, d_ref, "sgdet")
    assert rel_err(a[sel], ref) < FP32_TOL
    assert np.array_equal(a[sel].argmax(1), ref.argmax(1))


def test_relation_head_dropin():
    """veto_b200.relation_head.ROIRelationHead (relation_head.py:27-248), the caller of the path: at test time it
    reproduces the golden post-processed output of the reference's own head pipeline from raw proposals (labels only:
    the head overloads the PredCls fields itself); in training it samples, runs the fused step and back-propagates."""
    from veto_b200.relation_head import build_roi_relation_head
    from veto_b200.structures import BoxList
    name = "cfg1_predcls_vg"
    c, batch, state, g = CASES[name], case_batch(CASES[name]), case_state(CASES[name]), load_golden(name)
    cfg = H.make_cfg(precision="bf16x3")
    head = build_roi_relation_head(cfg, 256)
    head.predictor.load_state_dict(synth.to_torch_state(state), strict=True)
    head = head.to(DEV).eval()
    feats, depth = H.device_features(batch, DEV)

    def raw_proposals():
        out = []
        for i in range(batch["B"]):
            bl = BoxList(_t(batch["boxes"][i]), (batch["W"], batch["H"]), "xyxy")
            bl.add_field("labels", _t(batch["labels"][i]))
            out.append(bl)
        return out

    with torch.no_grad():
        roi, result, losses = head(feats, raw_proposals(), depth_features=depth)
    assert losses == {} and tuple(roi.shape) == (sum(batch["n_boxes"]), 256, 8, 8)
    pp = np.concatenate([H.np_(r.get_field("rel_pair_idxs")) for r in result])
    labels = np.concatenate([H.np_(r.get_field("pred_rel_labels")) for r in result])
    scores = np.concatenate([H.np_(r.get_field("pred_rel_scores")) for r in result])[:, 1:].max(1)
    assert np.allclose(scores, g["post_scores"], rtol=5e-4)
    s = g["post_scores"].astype(np.float64)
    gap = np.full(len(s), np.inf)
    gap[1:] = np.minimum(gap[1:], s[:-1] - s[1:])
    gap[:-1] = np.minimum(gap[:-1], s[:-1] - s[1:])
    clear = gap > 2e-3 * s
    assert np.array_equal(pp[clear], g["post_pairs"][clear]) and np.array_equal(labels[clear], g["post_labels"][clear])
    # training through the head: sampler -> extractor -> predictor -> losses
    head.train()
    props = raw_proposals()
    mats = synth.make_relation_matrices(3, batch["n_boxes"], 51, 10)
    targets = []
    for pbl, m in zip(props, mats):
        t = BoxList(pbl.bbox, pbl.size, "xyxy")
        t.add_field("relation", _t(m))
        targets.append(t)
    depth_t = depth.clone().requires_grad_(True)
    roi, props_out, losses = head(feats, props, depth_features=depth_t, targets=targets)
    assert set(losses) == {"rel_loss"} and props_out[0].has_field("locating_match")
    losses["rel_loss"].backward()
    grads = [p.grad for p in head.predictor.parameters() if p.grad is not None]
    assert len(grads) >= 80 and all(bool(torch.isfinite(gr).all()) for gr in grads) and depth_t.grad is not None
    assert float(losses["rel_loss"].detach()) > 0
    with pytest.raises(NotImplementedError):
        bad = H.make_cfg()
        bad.MODEL.ROI_RELATION_HEAD.PREDICTOR = "MotifPredictor"
        build_roi_relation_head(bad, 256)


@pytest.mark.parametrize("mode", ["predcls", "sgdet"])
def test_image_sharding_is_exact(mode):
    """BASELINE.json configs[4] (per-image sharded PredCls + SGDet sweep): images are independent, so running the
    pair-count-balanced shards of a mixed batch one by one (what ranks 0..3 of a 4-GPU job would each run) reproduces
    the unsharded batch BITWISE — logits, pairs and ROI features — in the headline bf16x3 mode."""
    from veto_b200 import distributed as vdist
    n_boxes = [20, 80, 20, 33, 80, 5, 20, 1, 47]
    batch = synth.make_batch(51, n_boxes, H=320, W=416, mode=mode)
    state = synth.predictor_state(13)
    cfg = H.make_cfg(mode=mode, max_pairs=8192, precision="bf16x3")
    full = H.run_head(cfg, state, batch, DEV, post=False)
    shards = vdist.shard_images(n_boxes, 4, 8192)
    assert sorted(i for s in shards for i in s) == list(range(len(n_boxes)))
    loads = [sum(vdist.pair_count(n_boxes[i], 8192) for i in s) for s in shards]
    biggest = max(vdist.pair_count(n, 8192) for n in n_boxes)            # LPT balance by pair count, not image count
    assert max(loads) <= max(biggest, 4.0 / 3.0 * sum(loads) / 4)
    for shard in shards:
        sub = dict(batch, B=len(shard), n_boxes=[n_boxes[i] for i in shard])
        for k in ("boxes", "labels", "predict_logits", "pred_labels", "pred_scores"):
            if k in batch:
                sub[k] = [batch[k][i] for i in shard]
        sub["feats"] = [f[shard] for f in batch["feats"]]
        sub["depth"] = batch["depth"][shard]
        part = H.run_head(cfg, state, sub, DEV, post=False)
        box_off = np.concatenate([[0], np.cumsum(n_boxes)])
        for j, i in enumerate(shard):
            assert torch.equal(part["pairs"][j], full["pairs"][i])
            assert torch.equal(part["rel_dists"][j], full["rel_dists"][i]), (mode, i)
        rows = np.concatenate([np.arange(box_off[i], box_off[i + 1]) for i in shard])
        assert torch.equal(part["x2d"], full["x2d"][rows]) and torch.equal(part["d2d"], full["d2d"][rows])


def test_no_cpu_fallback_and_errors():
    with pytest.raises(RuntimeError):
        ops.roi_align_forward(torch.zeros(1, 1, 4, 4), torch.zeros(1, 5), 1.0, 2, 2, 2)      # CPU tensors
    sg = H.make_cfg(predictor="VETOPredictor_MEET", mode="sgdet")
    meet = H.build_predictor(sg, synth.meet_state(1, 151, synth.GROUP_SPLITS[("VG", "divide4")]), DEV)
    batch = synth.make_batch(3, [3], H=320, W=416, mode="sgdet")
    with pytest.raises(KeyError):      # MEET sgdet test reads the detector's boxes_per_cls field like the reference (:3778)
        meet(H.boxlists(batch, DEV, 151), [torch.zeros((1, 2), dtype=torch.long, device=DEV)], None, None)
    with pytest.raises(ValueError):
        ops.make_config(151, 51, precision="fp8")
    cfg = ops.make_config(151, 51, "fp32", dim=512)
    import ctypes
    assert L.load().veto_packed_bytes(ctypes.byref(cfg)) == 0
    assert b"unsupported architecture" in L.load().veto_last_error()


@pytest.mark.parametrize("n_seq", [1, 6, 37, 300])
def test_attention_tcgen05_kernel(n_seq):
    """The tcgen05 attention kernel (six sequences per 128-row tile, block-diagonal softmax): bf16x3 split within
    3e-5 of fp64, single-pass bf16 within 1e-2."""
    g = torch.Generator().manual_seed(4)
    qkv = torch.randn(n_seq * 19, 1728, generator=g)
    q, k, v = [t.reshape(n_seq, 19, 6, 96).permute(0, 2, 1, 3).double() for t in qkv.chunk(3, -1)]
    att = torch.softmax(q @ k.transpose(-1, -2) * (96 ** -0.5), -1) @ v
    ref = att.permute(0, 2, 1, 3).reshape(n_seq * 19, 576).numpy()
    assert rel_err(H.np_(ops.test_attention_tc(_t(qkv.numpy()), True)), ref) < 3e-5
    assert rel_err(H.np_(ops.test_attention_tc(_t(qkv.numpy()), False)), ref) < 1e-2


def test_layout_and_schedule_switches_are_bit_identical():
    """The switches of the library that only change WHERE or by WHICH CTA something is computed — the opt-in 4-CTA-cluster
    variant of the encoder GEMMs (W tile multicast between two row tiles), the (sequence, head) item layout of q / k / v,
    256-wide column tiles — must give the same logits bit for bit as the defaults: same products, same K order.  The switches are read once per process, so every arm runs in a
    child process."""
    import os
    import subprocess
    import sys

    code = ("import numpy as np, sys; from tests import test_gpu_parity as T;"
            "out = [r.detach().cpu().numpy().ravel() for n in ('cfg1_predcls_vg', 'sgdet_cap', 'ragged_predcls')"
            " for p in ('f16c8', 'bf16x3') for r in T._run_predictor(n, p)[6][1]];"
            "np.save(sys.argv[1], np.concatenate(out))")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    arms = [{}, {"VETO_GEMM_CLUSTER4": "2"}, {"VETO_QKV_ITEM_LAYOUT": "0"}, {"VETO_GEMM_BN256": "0"}]
    outs = []
    for k, arm in enumerate(arms):
        path = os.path.join("/tmp", "veto_switch_%d_%d.npy" % (k, os.getpid()))
        env = dict(os.environ, PYTHONPATH=root, **arm)
        r = subprocess.run([sys.executable, "-c", code, path], cwd=root, env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, (arm, r.stderr[-2000:])
        outs.append(np.load(path))
        os.remove(path)
    assert outs[0].size > 0
    for arm, o in zip(arms[1:], outs[1:]):
        assert o.shape == outs[0].shape and np.array_equal(o, outs[0]), arm
