"""GPU diagnostics (not collected by pytest): staged checks of libveto_b200.so with verbose numeric reports.

  python tests/gpu_diag.py <stage> [...]     stages: simt tc rows pairs gather head:<precision> speed:<precision>

Each stage runs in its own process from tools/gpu_check.sh so that a trapped kernel cannot poison the next.
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from veto_b200 import lib as L  # noqa: E402
from veto_b200 import ops, synth  # noqa: E402

DEV = torch.device("cuda:0")


def report(name, **kw):
    print(json.dumps(dict(stage=name, **{k: (float(v) if isinstance(v, (np.floating, float)) else v) for k, v in kw.items()})),
          flush=True)


def relerr(a, ref):
    a, ref = a.double(), ref.double()
    return float((a - ref).abs().max() / ref.abs().max().clamp_min(1e-30))


def stage_simt():
    g = torch.Generator(device="cpu").manual_seed(0)
    for (M, N, K) in [(300, 51, 576), (1000, 1152, 200), (77, 128, 128), (4864, 576, 1152)]:
        a = torch.randn(M, K, generator=g).to(DEV)
        w = torch.randn(N, K, generator=g).to(DEV)
        b = torch.randn(N, generator=g).to(DEV)
        r = torch.randn(M, N, generator=g).to(DEV)
        ref = a.double() @ w.double().T
        report("simt", shape=[M, N, K], err_plain=relerr(ops.test_gemm(a, w), ref),
               err_bias_relu=relerr(ops.test_gemm(a, w, bias=b, act=1), torch.relu(ref + b.double())),
               err_gelu_res=relerr(ops.test_gemm(a, w, bias=b, residual=r, act=2),
                                   torch.nn.functional.gelu(ref + b.double()) + r.double()))


def stage_tc():
    g = torch.Generator(device="cpu").manual_seed(1)
    shapes = [(128, 192, 64), (128, 128, 64), (128, 192, 576), (256, 576, 1152), (1000, 1728, 576), (320, 1024, 1024),
              (320, 128, 1024), (4864, 1152, 576), (19456, 1728, 576)]
    for (M, N, K) in shapes:
        a = torch.randn(M, K, generator=g).to(DEV)
        w = (torch.randn(N, K, generator=g) / K ** 0.5).to(DEV)
        b = torch.randn(N, generator=g).to(DEV)
        ref32 = a.double() @ w.double().T
        a16, w16 = a.bfloat16().double(), w.bfloat16().double()
        ref16 = a16 @ w16.T
        for prec, ref in (("bf16", ref16), ("bf16x3", ref32)):
            try:
                c = ops.test_gemm(a, w, precision=prec)
                torch.cuda.synchronize()
                e = relerr(c, ref)
                info = dict(shape=[M, N, K], precision=prec, err=e)
                if e > 1e-3:
                    d = (c.double() - ref).abs()
                    info["err_by_rowblock32"] = [float(x) for x in d[:128].reshape(4, 32, -1).amax((1, 2))]
                    info["err_by_colblock32"] = [float(x) for x in d[:, :min(N, 192)].reshape(M, -1, 32).amax((0, 2))]
                    info["c00"] = [float(x) for x in c[0, :8]]
                    info["ref00"] = [float(x) for x in ref[0, :8]]
                    info["ratio_mean"] = float((c.double() / ref).median())
                c2 = ops.test_gemm(a, w, bias=b, act=2, precision=prec)
                info["err_bias_gelu"] = relerr(c2, torch.nn.functional.gelu(ref + b.double()))
                report("tc", **info)
            except Exception as ex:  # noqa: BLE001
                report("tc", shape=[M, N, K], precision=prec, error=str(ex)[:400])
                return


def stage_rows():
    g = torch.Generator(device="cpu").manual_seed(2)
    x = (3 * torch.randn(1000, 576, generator=g) + 0.5).to(DEV)
    w = torch.randn(576, generator=g).to(DEV)
    b = torch.randn(576, generator=g).to(DEV)
    ref = torch.nn.functional.layer_norm(x.double(), (576,), w.double(), b.double(), 1e-5)
    report("layernorm", err=relerr(ops.test_layernorm(x, w, b), ref))
    n_seq = 37
    qkv = torch.randn(n_seq * 19, 1728, generator=g).to(DEV)
    q, k, v = [t.reshape(n_seq, 19, 6, 96).permute(0, 2, 1, 3).double() for t in qkv.chunk(3, -1)]
    att = torch.softmax(q @ k.transpose(-1, -2) * (96 ** -0.5), -1) @ v
    ref = att.permute(0, 2, 1, 3).reshape(n_seq * 19, 576)
    report("attention", err=relerr(ops.test_attention(qkv), ref))
    for n_seq in (37, 6, 1, 300):
        qkv = torch.randn(n_seq * 19, 1728, generator=g).to(DEV)
        q, k, v = [t.reshape(n_seq, 19, 6, 96).permute(0, 2, 1, 3).double() for t in qkv.chunk(3, -1)]
        att = torch.softmax(q @ k.transpose(-1, -2) * (96 ** -0.5), -1) @ v
        ref = att.permute(0, 2, 1, 3).reshape(n_seq * 19, 576)
        for split in (True, False):
            out = ops.test_attention_tc(qkv, split)
            torch.cuda.synchronize()
            d = (out.double() - ref).abs()
            report("attention_tc", n_seq=n_seq, split=split, err=relerr(out, ref), nan=int(torch.isnan(out).sum()),
                   err_by_token=[float(x) for x in d.reshape(n_seq, 19, 576).amax((0, 2))][:19] if relerr(out, ref) > 1e-2 else None,
                   err_by_head=[float(x) for x in d.reshape(-1, 6, 96).amax((0, 2))] if relerr(out, ref) > 1e-2 else None)


def _case(name):
    from tests.cases import CASES, case_batch, case_state, load_golden
    c = CASES[name]
    return c, case_batch(c), case_state(c), load_golden(name)


def stage_pairs():
    from oracle import veto_oracle as O
    for name in ("cfg1_predcls_vg", "ragged_predcls", "sgdet_cap", "sgdet_overlap"):
        c, batch, _, g = _case(name)
        boxes = torch.from_numpy(np.concatenate(batch["boxes"])).to(DEV)
        scores = torch.from_numpy(np.concatenate(batch["pred_scores"])).to(DEV) if "pred_scores" in batch else None
        pairs = ops.enumerate_pairs(batch["n_boxes"], DEV, c.get("max_pairs", 2048), boxes=boxes, scores=scores,
                                    require_overlap=c.get("require_overlap", False) and c["mode"] == "sgdet")
        ref = O.prepare_test_pairs(batch["n_boxes"], c.get("max_pairs", 2048), scores=batch.get("pred_scores"),
                                   boxes=batch["boxes"],
                                   require_overlap=c.get("require_overlap", False) and c["mode"] == "sgdet")
        ok = all(np.array_equal(p.cpu().numpy(), r) for p, r in zip(pairs, ref))
        report("pairs", case=name, counts=[int(p.shape[0]) for p in pairs], equal_oracle=bool(ok),
               equal_golden=bool(np.array_equal(np.concatenate([p.cpu().numpy() for p in pairs]), g["pairs"])))


def stage_gather():
    from oracle import veto_oracle as O
    for name in ("cfg1_predcls_vg", "ragged_predcls"):
        c, batch, _, g = _case(name)
        feats = [torch.from_numpy(f).to(DEV) for f in batch["feats"]]
        depth = torch.from_numpy(batch["depth"]).to(DEV)
        boxes = torch.from_numpy(np.concatenate(batch["boxes"])).to(DEV)
        x2d, d2d, lv = ops.roi_gather(feats, depth, boxes, batch["n_boxes"], synth.POOLER_SCALES, synth.DEPTH_SCALE,
                                      return_levels=True)
        rx, rd = O.pooler_forward(batch["feats"], batch["depth"], batch["boxes"])
        x, d = x2d.cpu().numpy(), d2d.cpu().numpy()
        report("gather", case=name, rgb_bit_exact=bool(np.array_equal(x, rx)), depth_bit_exact=bool(np.array_equal(d, rd)),
               rgb_maxdiff=float(np.abs(x - rx).max()), depth_maxdiff=float(np.abs(d - rd).max()),
               levels_equal=bool(np.array_equal(lv.cpu().numpy(), O.level_map(batch["boxes"]))),
               golden_sub=bool(np.array_equal(x[:, ::16], g["x2d_sub"]) and np.array_equal(d[:, ::16], g["d2d_sub"])))
        # the one-for-one _C.roi_align_forward replacement on one level
        rois = torch.from_numpy(O.rois_format(batch["boxes"])).to(DEV)
        y = ops.roi_align_forward(feats[1], rois, 0.125, 8, 8, 2).cpu().numpy()
        report("roi_align", case=name, bit_exact=bool(np.array_equal(y, O.roi_align(batch["feats"][1], O.rois_format(batch["boxes"]), 0.125))))


def stage_head(precision):
    from tests import harness as H
    for name in ("cfg1_predcls_vg", "ragged_predcls", "sgdet_cap", "meet_gqa"):
        c, batch, state, g = _case(name)
        cfg = H.make_cfg(c["predictor"], c["mode"], c["dataset"], c.get("max_pairs", 2048), c.get("require_overlap", False),
                         precision)
        t0 = time.time()
        out = H.run_head(cfg, state, batch, DEV)
        torch.cuda.synchronize()
        if isinstance(out["rel_dists"], dict):
            errs = {k: relerr(v, torch.from_numpy(g["logits_" + k]).to(DEV)) for k, v in out["rel_dists"].items()}
            report("head", case=name, precision=precision, err=max(errs.values()), sec=time.time() - t0)
            continue
        logits = torch.cat(list(out["rel_dists"]))
        ref = torch.from_numpy(g["logits"]).to(DEV)
        same_pairs = bool(np.array_equal(np.concatenate([H.np_(p) for p in out["pairs"]]), g["pairs"]))
        info = dict(case=name, precision=precision, same_pairs=same_pairs, sec=time.time() - t0)
        if same_pairs:
            info["err"] = relerr(logits, ref)
            info["argmax_equal"] = bool((logits[:, 1:].argmax(1) == ref[:, 1:].argmax(1)).all())
            info["n_labels"] = int(ref[:, 1:].argmax(1).unique().numel())
        report("head", **info)


def stage_speed(precision):
    """Throughput of the relation head alone (inputs resident) on an SGDet-shaped batch."""
    import ctypes
    n_img, n_box = (int(os.environ.get("DIAG_IMAGES", 4)), int(os.environ.get("DIAG_BOXES", 80)))
    batch = synth.make_batch(7, [n_box] * n_img, features=False)
    state = synth.predictor_state(11)
    from tests import harness as H
    cfg = H.make_cfg(precision=precision, max_pairs=8192)
    pred = H.build_predictor(cfg, state, DEV)
    bls = H.boxlists(batch, DEV, 151)
    N = n_img * n_box
    x2d = torch.randn(N, 256, 8, 8, device=DEV)
    d2d = torch.relu(torch.randn(N, 256, 8, 8, device=DEV))
    pairs = ops.enumerate_pairs(batch["n_boxes"], DEV, 8192)
    R = sum(p.shape[0] for p in pairs)
    for chunk in [int(v) for v in os.environ.get("DIAG_CHUNKS", "256,512,1024,2048").split(",")]:
        pred.chunk_pairs = chunk
        with torch.no_grad():
            for _ in range(2):
                pred(bls, pairs, None, None, roi_features=x2d, roi_depth_features=d2d)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                pred(bls, pairs, None, None, roi_features=x2d, roi_depth_features=d2d)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        report("speed", precision=precision, pairs=R, chunk=chunk, ms=ms, pairs_per_s=R / ms * 1e3,
               tflops_ref=R * 648.7e6 / ms / 1e9, launches=ops.last_launch_count)


def stage_gemm(precision):
    """GEMM micro-benchmark at the encoder's shapes (one default chunk = 37886 token rows): TFLOP/s per epilogue."""
    M = int(os.environ.get("DIAG_M", 37886))
    passes = 3 if precision == "bf16x3" else 1
    g = torch.Generator(device="cpu").manual_seed(3)
    for (N, K, name) in [(1728, 576, "qkv"), (576, 576, "out"), (1152, 576, "ff1"), (576, 1152, "ff2")]:
        a = torch.randn(M, K, generator=g).to(DEV)
        w = (torch.randn(N, K, generator=g) / K ** 0.5).to(DEV)
        b = torch.randn(N, generator=g).to(DEV)
        r = torch.randn(M, N, generator=g).to(DEV)
        c = torch.empty(M, N, device=DEV)
        scratch = torch.empty(4 * (M * K + N * K) + 4 * M * N + 256, dtype=torch.uint8, device=DEV)
        variants = {"plain_f32": (None, None, 0), "bias_f32": (b, None, 0), "bias_res_f32": (b, r, 0),
                    "bias_gelu_f32": (b, None, 2), "bias_gelu_split": (b, None, 2 | 0x100), "plain_split": (None, None, 0x100)}
        out = {}
        for vn, (bias, res, act) in variants.items():
            ops.test_gemm(a, w, bias=bias, residual=res, act=act, precision=precision, scratch=scratch, c=c)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                ops.test_gemm(a, w, bias=bias, residual=res, act=act | 0x200, precision=precision, scratch=scratch, c=c)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            out[vn] = round(2.0 * M * N * K * passes / ms / 1e9, 1)
        report("gemm", precision=precision, shape=[M, N, K], gemm=name, executed_tflops=out)


def stage_stages(precision):
    """Per-stage device time (CUDA events around every library launch) of the relation head on an SGDet-shaped batch."""
    n_img, n_box = (int(os.environ.get("DIAG_IMAGES", 8)), int(os.environ.get("DIAG_BOXES", 80)))
    batch = synth.make_batch(7, [n_box] * n_img, features=False)
    from tests import harness as H
    cfg = H.make_cfg(precision=precision, max_pairs=8192, chunk_pairs=int(os.environ.get("DIAG_CHUNK", 0)))
    pred = H.build_predictor(cfg, synth.predictor_state(11), DEV)
    bls = H.boxlists(batch, DEV, 151)
    N = n_img * n_box
    x2d = torch.randn(N, 256, 8, 8, device=DEV)
    d2d = torch.relu(torch.randn(N, 256, 8, 8, device=DEV))
    pairs = ops.enumerate_pairs(batch["n_boxes"], DEV, 8192)
    R = sum(p.shape[0] for p in pairs)
    with torch.no_grad():
        for _ in range(2):
            pred(bls, pairs, None, None, roi_features=x2d, roi_depth_features=d2d)
        torch.cuda.synchronize()
        with ops.StageTimer() as st:
            pred(bls, pairs, None, None, roi_features=x2d, roi_depth_features=d2d)
    tot = sum(st.ms.values())
    report("stages", precision=precision, pairs=R, total_ms=tot, us_per_pair=tot / R * 1e3,
           ms={k: round(v, 3) for k, v in st.ms.items()}, launches=st.launches)


if __name__ == "__main__":
    L.require_device()
    for arg in sys.argv[1:]:
        name, _, opt = arg.partition(":")
        fn = globals()["stage_" + name]
        try:
            fn(opt) if opt else fn()
        except Exception as ex:  # noqa: BLE001
            import traceback
            traceback.print_exc()
            report(name, fatal=str(ex)[:500])
            sys.exit(1)
