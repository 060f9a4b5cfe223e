"""f3, the depth backbone (ResNetDepth, backbone/resnet_depth.py:11-47).

CPU: the oracle restatement (oracle/depth_port.py) against the fixture the UNMODIFIED reference module produced
(tests/golden/depth_backbone.npz; generator: tests/golden/make_golden.py run_depth_backbone), and the host mirror's
state-dict contract.  GPU (-m gpu): the CUDA forward / backward through the C ABI against the oracle and the fixture.
"""
import hashlib

import numpy as np
import pytest
import torch

from oracle import depth_port as P
from tests.cases import load_golden
from tests.train_util import check_against_golden, grad_error

# tolerances, relative to the largest magnitude of the compared tensor: the fp32-grade modes are bounded by fp32
# accumulation-order noise amplified through 15 batch-statistics BatchNorms; single-pass bf16 is reported, loosely bounded
OUT_TOL = {"fp32": 2e-4, "bf16x3": 2e-4, "bf16": 8e-2}
# (bf16x3 keeps 16 mantissa bits per operand: a handful of ReLU / arg-max decisions on near-ties fall the other way than in
# the reference's fp32 run, which moves single gradient entries by a few 1e-3 of the tensor's scale — see the sensitivity
# measurement in test_gpu_matches_oracle_full_gradients; the tensor norms agree to 1e-4)
GRAD_TOL = {"fp32": 2e-3, "bf16x3": 5e-3}


def _digest(arrs):
    h = hashlib.sha256()
    for a in arrs:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def _fixture():
    g = load_golden("depth_backbone")
    sd = P.synth_state(0)
    depth = P.synth_depth(*[int(g["shape"][i]) for i in (0, 2, 3)])
    assert _digest([depth]) == str(g["input_digest"]) and _digest([sd[k] for k in sorted(sd)]) == str(g["weight_digest"])
    grad_out = np.random.RandomState(5).standard_normal(g["train_out"].shape).astype(np.float32)
    assert _digest([grad_out]) == str(g["grad_out_digest"])
    return g, sd, depth, grad_out


def test_oracle_matches_reference_fixture():
    g, sd, depth, grad_out = _fixture()
    assert P.state_keys() == [str(k) for k in g["keys"]]
    ev = P.eval_forward(sd, depth)
    assert grad_error(ev, g["eval_out"]) < 1e-5
    out, grads, stats = P.train_step(sd, depth, grad_out)
    assert grad_error(out, g["train_out"]) < 1e-5
    check_against_golden(grads, g, 1e-3)
    for k, v in stats.items():
        np.testing.assert_allclose(v, g["buf/" + k], rtol=1e-5, atol=1e-6)


def test_host_mirror_state_dict_contract():
    """Same keys, order and shapes as the reference module; registered under the reference's registry key."""
    import veto_b200
    from veto_b200 import depth_backbone as D
    from veto_b200 import registry
    veto_b200.load_modules()
    g = load_golden("depth_backbone")
    model = registry.BACKBONES["R-18-C4"](None, True)
    assert model.out_channels == int(g["out_channels"]) == 256
    sd = model.state_dict()
    assert list(sd.keys()) == [str(k) for k in g["keys"]]
    ref = P.synth_state(0)
    for k, v in sd.items():
        assert tuple(v.shape) == tuple(np.asarray(ref[k]).shape), k
    model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in ref.items()}, strict=True)
    # initialisation of resnet_depth.py:27-28 / torchvision ResNet.__init__: He fan-out, BatchNorm (1, 0)
    fresh = D.ResNetDepth()
    w = fresh.layer2[0].conv1.weight.detach().numpy()
    assert abs(w.std() / np.sqrt(2.0 / (9 * 128)) - 1) < 0.05
    assert float(fresh.bn1.weight.min()) == 1.0 and float(fresh.bn1.bias.abs().max()) == 0.0
    assert D.build_backbone(None, depth_backbone=True).out_channels == 256
    with pytest.raises(NotImplementedError):
        D.build_backbone(None, depth_backbone=False)


def test_output_size_and_workspace_host_side():
    """Host-only entry points (no GPU needed): the stride-16 output size follows torch's floor rule for every image size,
    and the workspace grows with the batch and with training mode."""
    from veto_b200 import lib as L
    from veto_b200 import ops
    lib = L.load()
    rng = np.random.RandomState(0)
    for h, w in [(16, 16), (17, 31), (592, 800), (600, 1000), (608, 1008), (1333, 800)] + [tuple(rng.randint(16, 1400, 2)) for _ in range(40)]:
        assert ops.depth_backbone_out_size(int(h), int(w)) == P.out_size(int(h), int(w)), (h, w)
    x = torch.zeros(1, 1, 45, 77)
    assert tuple(P.eval_forward(P.synth_state(0), x.numpy()).shape[2:]) == ops.depth_backbone_out_size(45, 77)
    sizes = [lib.veto_depth_backbone_workspace_bytes(1, b, 592, 800, t) for b, t in ((1, 0), (1, 1), (12, 0), (12, 1))]
    assert 0 < sizes[0] < sizes[1] < sizes[3] and sizes[0] < sizes[2] < sizes[3]
    assert lib.veto_depth_backbone_workspace_bytes(7, 1, 64, 64, 0) == 0          # bad precision code


def test_product_path_needs_the_device():
    from veto_b200 import depth_backbone as D
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    with pytest.raises(RuntimeError):
        D.ResNetDepth().eval()(torch.zeros(1, 1, 32, 32))


# ---------------------------------------------------------------------------------------------------------------
def _model(sd, precision, device="cuda"):
    from veto_b200 import depth_backbone as D
    body = D.ResNetDepth(precision)
    model = torch.nn.Sequential()
    model.add_module("body", body)
    model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=True)
    return model.to(device)


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "bf16x3", "bf16"])
def test_gpu_matches_reference_fixture(precision):
    g, sd, depth, grad_out = _fixture()
    model = _model(sd, precision).eval()
    x = torch.from_numpy(depth).cuda()
    ev = model(x)
    assert not ev.requires_grad
    assert grad_error(ev.cpu().numpy(), g["eval_out"]) < OUT_TOL[precision]
    model.train()
    y = model(x)
    assert grad_error(y.detach().cpu().numpy(), g["train_out"]) < OUT_TOL[precision]
    y.backward(torch.from_numpy(grad_out).cuda())
    grads = {k: p.grad.cpu().numpy() for k, p in model.named_parameters()}
    if precision == "bf16":
        # single-pass bf16 through 15 batch-statistics BatchNorms: individual small gradient entries are noise-level,
        # the tensors as a whole are not — bound the norm of every gradient tensor
        for k, v in grads.items():
            norm_ref = float(g["gstat/" + k][0])
            assert abs(np.sqrt((v.astype(np.float64) ** 2).sum()) - norm_ref) <= 0.05 * norm_ref, k
    else:
        worst = check_against_golden(grads, g, GRAD_TOL[precision])
        assert len(worst) == 45
    bufs = dict(model.named_buffers())
    for k in g.files:
        if k.startswith("buf/"):
            tol = 1e-4 if precision != "bf16" else 5e-2
            np.testing.assert_allclose(bufs[k[4:]].cpu().numpy(), g[k], rtol=tol, atol=tol)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(1, 64, 64), (3, 97, 131), (2, 160, 224)])
def test_gpu_matches_oracle_full_gradients(shape):
    """Every gradient element against the fp64 oracle, odd image sizes (ragged borders in every strided stage)."""
    B, H, W = shape
    sd = P.synth_state(3)
    depth = P.synth_depth(B, H, W, seed=7)
    from veto_b200 import ops
    oh, ow = ops.depth_backbone_out_size(H, W)
    grad_out = np.random.RandomState(11).standard_normal((B, 256, oh, ow)).astype(np.float32)
    out64, grads64, stats64 = P.train_step(sd, depth.astype(np.float64), grad_out.astype(np.float64), dtype=torch.float64)
    assert out64.shape == grad_out.shape
    out32, grads32, _ = P.train_step(sd, depth, grad_out)                   # the fp32 CPU path's own distance to fp64
    # The gradient of this network is discontinuous in the weights (ReLU masks, max-pool arg-max, batch statistics
    # over few positions): the oracle ITSELF moves by `sens` when its conv weights move by 1e-5 relative, which is
    # the operand resolution of the bf16x3 scheme (2^-17 per operand).  Each mode is held to the strict bound or to a
    # multiple of that sensitivity, whichever is larger (test_gpu_matches_reference_fixture holds both to the strict one).
    rng = np.random.RandomState(1)
    sd_p = {k: (v * (1 + 1e-5 * rng.standard_normal(v.shape))).astype(np.float32) if (v.ndim == 4) else v for k, v in sd.items()}
    _, grads_p, _ = P.train_step(sd_p, depth.astype(np.float64), grad_out.astype(np.float64), dtype=torch.float64)
    sens = max(grad_error(grads_p[k], grads64[k]) for k in grads64)
    for precision in ("fp32", "bf16x3"):
        model = _model(sd, precision).train()
        y = model(torch.from_numpy(depth).cuda())
        y.backward(torch.from_numpy(grad_out).cuda())
        assert grad_error(y.detach().cpu().numpy(), out64) < OUT_TOL[precision]
        # (fp32 on the GPU sums in another order than the CPU, a 1e-6 perturbation: a tenth of the sensitivity scale)
        bound = max(GRAD_TOL[precision], (0.3 if precision == "fp32" else 3.0) * sens)
        for k, p in model.named_parameters():
            e = grad_error(p.grad.cpu().numpy(), grads64[k])
            floor = grad_error(grads32[k], grads64[k])
            assert e < max(bound, 20 * floor), f"{precision} {k}: {e:.2e} (fp32 CPU: {floor:.2e}, sensitivity {sens:.2e})"
        for k, b in model.named_buffers():
            if "running_" in k:
                np.testing.assert_allclose(b.cpu().numpy(), stats64[k], rtol=1e-4, atol=1e-5)
            else:
                assert int(b) == 1


@pytest.mark.gpu
def test_gpu_full_size_training_step_properties():
    """BASELINE-size batch (the sizes the oracle cannot finish in seconds): shape, finiteness, BatchNorm invariants
    (the per-channel sum of d(conv output) is zero => the conv bias-free weight gradients of a constant input shift
    vanish is not observable, so check what is: running statistics moved towards the batch statistics, a second
    backward raises, eval after train is deterministic)."""
    sd = P.synth_state(1)
    model = _model(sd, "bf16x3").train()
    depth = torch.from_numpy(P.synth_depth(2, 608, 1008, seed=3)).cuda()
    y = model(depth)
    assert tuple(y.shape) == (2, 256, 38, 63) and bool(torch.isfinite(y).all()) and float(y.detach().min()) >= 0.0
    y.sum().backward(retain_graph=True)
    for k, p in model.named_parameters():
        assert p.grad is not None and bool(torch.isfinite(p.grad).all()), k
    # d/d(beta) of sum(relu(..)) for the last BatchNorm = number of active outputs per channel
    last = model.body.layer3[1].bn2
    active = (y > 0).sum(dim=(0, 2, 3)).float()
    assert torch.allclose(last.bias.grad, active, rtol=1e-5, atol=0.5)
    with pytest.raises(RuntimeError):
        y.sum().backward()
    model.eval()
    a, b = model(depth), model(depth)
    assert torch.equal(a, b)


@pytest.mark.gpu
def test_gpu_chain_depth_backbone_into_relation_head():
    """End to end as relation_train_net.py trains it: depth image -> depth backbone (train()) -> VETOFeatureExtractor ->
    VETOPredictor train() -> rel_loss.backward() -> gradients of the BACKBONE's parameters, i.e. through both C calls and
    the ROIAlign backward, against the same chain on the CPU oracle (depth_port -> torchvision roi_align -> torch_port)."""
    from oracle import torch_port as TP
    from tests import harness as H
    from tests.cases import TRAIN_CASES
    from tests.test_gpu_train import DEV, _set_dropout
    from tests.train_util import train_case_inputs
    from veto_b200 import config as vcfg
    from veto_b200 import registry, synth
    c = TRAIN_CASES["train_predcls"]
    batch, sd, pairs, labels = train_case_inputs(c)
    B, _, hd, wd = batch["depth"].shape
    dsd, dimg = P.synth_state(2), P.synth_depth(B, 16 * hd, 16 * wd, seed=4)
    assert P.out_size(16 * hd, 16 * wd) == (hd, wd)
    # ---- CPU oracle chain
    st = {k: torch.from_numpy(np.asarray(v)).clone().requires_grad_(v.dtype.kind == "f" and "running_" not in k)
          for k, v in dsd.items()}
    dmap = P.forward(st, torch.from_numpy(dimg), True)
    boxes = [torch.from_numpy(b) for b in batch["boxes"]]
    x2d, d2d = TP.pooler_forward([torch.from_numpy(f) for f in batch["feats"]], dmap, boxes)
    tsd = TP.to_torch(sd)
    loss_ref, _, g_d2d, *_ = TP.train_step(
        tsd, boxes, [torch.from_numpy(p) for p in pairs], [torch.from_numpy(l) for l in labels], x2d.detach(), d2d.detach(),
        c["mode"], labels=[torch.from_numpy(l) for l in batch["labels"]], class_weight=tsd["criterion_loss_rel.weight"])
    d2d.backward(g_d2d)
    # ---- the product path.  A wiring error is O(1).  fp32 mode: accumulation-order noise only; bf16x3: its 16-bit operands
    # flip single near-tie ReLU / arg-max decisions, each moving the entries behind it by a few 1e-2 of the tensor's scale
    # (see the sensitivity measurement in test_gpu_matches_oracle_full_gradients).
    for precision, med_tol, max_tol in (("fp32", 5e-3, 1e-1), ("bf16x3", 3e-2, 3e-1)):
        cfg = H.make_cfg(predictor=c["predictor"], mode=c["mode"], dataset=c["dataset"], precision=precision)
        bls = H.boxlists(batch, DEV, vcfg.num_classes(cfg)[0])
        feats, _ = H.device_features(batch, DEV)
        backbone = _model(dsd, precision, DEV).train()
        fe = registry.make_roi_box_feature_extractor(cfg, 256, for_relation=True).to(DEV).train()
        pred = registry.make_roi_relation_predictor(cfg, 512)
        pred.load_state_dict(synth.to_torch_state(sd), strict=True)
        pred = pred.to(DEV).train()
        _set_dropout(pred, 0.0, 0.0, 0.0)
        depth_features = backbone(torch.from_numpy(dimg).to(DEV))
        assert depth_features.requires_grad and tuple(depth_features.shape) == (B, 256, hd, wd)
        x2d_g, d2d_g, _, _ = fe(feats, bls, depth_features=depth_features)
        loss = pred(bls, [torch.from_numpy(p).to(DEV) for p in pairs], [torch.from_numpy(l).to(DEV) for l in labels], None,
                    roi_features=x2d_g, roi_depth_features=d2d_g)[2]["rel_loss"]
        loss.backward()
        torch.cuda.synchronize()
        assert abs(float(loss.detach()) - float(loss_ref)) <= 1e-4 * abs(float(loss_ref))
        errs = {}
        for k, p in backbone.named_parameters():
            assert p.grad is not None and bool(torch.isfinite(p.grad).all()), k
            errs[k] = grad_error(p.grad.cpu().numpy(), st[k].grad.numpy())
        worst = sorted(errs.items(), key=lambda kv: -kv[1])[:4]
        assert np.median(list(errs.values())) < med_tol and max(errs.values()) < max_tol, (precision, worst)
