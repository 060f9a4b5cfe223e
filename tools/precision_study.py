#!/usr/bin/env python
"""Precision study of the tensor-core operand formats, by emulation on the CPU.

Every Linear of the encoder and the patch projections (the GEMMs the tcgen05 kernels run) is re-evaluated with its
operands rounded the way a candidate tensor-core mode would round them, products accumulated in fp32 (as TMEM does);
everything else (LayerNorm, attention core, small fp32 Linears, classifier) stays fp32.  The logits are compared with
the plain fp32 torch port on the golden-fixture inputs: max |diff| / max |ref| (the north_star metric, bar 1e-3), and
the number of argmax flips.

`units` = tensor-pipe cost in bf16-MMA equivalents (kind::f16 = 1, kind::tf32 = 2, kind::f8f6f4 = 0.5 per product).

    python tools/precision_study.py [case ...]        (build container or any CPU box; uses oracle/ = test infrastructure)
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import torch_port as TP  # noqa: E402
from tests.cases import CASES, case_batch, case_state  # noqa: E402


def r_bf16(x):
    return x.to(torch.bfloat16).float()


def r_fp16(x):
    return x.clamp(-65504, 65504).to(torch.float16).float()


def r_e4m3(x):
    return x.clamp(-448, 448).to(torch.float8_e4m3fn).float()


def r_tf32(x, trunc=False):
    i = x.contiguous().view(torch.int32)
    if trunc:
        i = i & ~0x1FFF
    else:  # round to nearest even on the 13 dropped bits
        i = (i + 0xFFF + ((i >> 13) & 1)) & ~0x1FFF
    return i.view(torch.float32)


def mm(a, w):
    return a @ w.t()


def make_mode(name):
    """-> (units, fn(a [M,K], w [N,K]) -> [M,N])"""
    if name == "fp32":
        return 0.0, mm
    if name == "bf16":
        return 1.0, lambda a, w: mm(r_bf16(a), r_bf16(w))
    if name == "fp16":
        return 1.0, lambda a, w: mm(r_fp16(a), r_fp16(w))
    if name == "tf32":
        return 2.0, lambda a, w: mm(r_tf32(a), r_tf32(w))
    if name == "tf32_trunc":
        return 2.0, lambda a, w: mm(r_tf32(a, True), r_tf32(w, True))
    if name in ("bf16x3", "fp16x3"):
        r = r_bf16 if name == "bf16x3" else r_fp16

        def f(a, w):
            ah, wh = r(a), r(w)
            al, wl = r(a - ah), r(w - wh)
            return mm(ah, wh) + mm(al, wh) + mm(ah, wl)
        return 3.0, f
    if name in ("bf16x2_a", "fp16x2_a"):     # activations split, weights single
        r = r_bf16 if name[0] == "b" else r_fp16

        def f(a, w):
            ah, wh = r(a), r(w)
            return mm(ah, wh) + mm(r(a - ah), wh)
        return 2.0, f
    if name in ("bf16x2_w", "fp16x2_w"):     # weights split, activations single
        r = r_bf16 if name[0] == "b" else r_fp16

        def f(a, w):
            ah, wh = r(a), r(w)
            return mm(ah, wh) + mm(ah, r(w - wh))
        return 2.0, f
    if name in ("fp16_fp8c", "bf16_fp8c"):
        # 16-bit main product + the two first-order corrections in fp8 e4m3 (kind::f8f6f4, twice the bf16 rate):
        # residuals scaled by 2^s into e4m3's range, the correction accumulator scaled back by 2^-s
        r, s = (r_fp16, 11) if name[0] == "f" else (r_bf16, 8)

        def f(a, w):
            ah, wh = r(a), r(w)
            al8, wl8 = r_e4m3((a - ah) * 2.0 ** s), r_e4m3((w - wh) * 2.0 ** s)
            return mm(ah, wh) + (mm(al8, r_e4m3(w)) + mm(r_e4m3(a), wl8)) * 2.0 ** -s
        return 2.0, f
    if name == "fp16_fp8c_wscaled":
        # as fp16_fp8c with per-output-row weight scaling of the e4m3 weight copy (weights are small: e4m3 subnormals)
        def f(a, w):
            ah, wh = r_fp16(a), r_fp16(w)
            ws = w.abs().amax(1, keepdim=True).clamp_min(1e-30)
            sc = 2.0 ** torch.floor(torch.log2(256.0 / ws))
            al8 = r_e4m3((a - ah) * 2.0 ** 11)
            wl8 = r_e4m3((w - wh) * sc * 2.0 ** 11)
            return mm(ah, wh) + (mm(al8, r_e4m3(w * sc)) + mm(r_e4m3(a), wl8)) * (2.0 ** -11 / sc.t())
        return 2.0, f
    raise KeyError(name)


MODES = ["bf16", "fp16", "tf32", "tf32_trunc", "bf16x2_a", "bf16x2_w", "fp16x2_a", "fp16x2_w", "bf16_fp8c", "fp16_fp8c",
         "fp16_fp8c_wscaled", "bf16x3", "fp16x3"]


class LinearPatch:
    """F.linear of oracle/torch_port.py replaced for the tensor-core GEMMs (in_features 576 / 1152 / 2048 and more than
    128 outputs: to_qkv, to_out, the feed-forward pair, proj_d; proj_v has 64 outputs but runs on the same kernel)."""

    def __init__(self, fn):
        self.fn = fn
        self.real = TP.F.linear

    def __call__(self, x, w, b=None):
        tc = w.dim() == 2 and w.shape[1] in (576, 1152, 2048) and not (w.shape[1] == 576 and w.shape[0] <= 128)
        if not tc:
            return self.real(x, w, b)
        y = self.fn(x.reshape(-1, x.shape[-1]), w).reshape(x.shape[:-1] + (w.shape[0],))
        return y if b is None else y + b

    def __enter__(self):
        TP.F.linear = self
        return self

    def __exit__(self, *exc):
        TP.F.linear = self.real
        return False


def run_case(name, modes):
    c = CASES[name]
    torch.set_num_threads(os.cpu_count() or 1)
    batch = case_batch(c)
    sd = TP.to_torch(case_state(c))
    feats = [torch.from_numpy(f) for f in batch["feats"]]
    depth = torch.from_numpy(batch["depth"])
    boxes = [torch.from_numpy(b) for b in batch["boxes"]]
    pairs = TP.prepare_test_pairs(batch["n_boxes"], c.get("max_pairs", 2048),
                                  [torch.from_numpy(s) for s in batch["pred_scores"]] if "pred_scores" in batch else None)
    x2d, d2d = TP.pooler_forward(feats, depth, boxes)
    kw = dict(labels=[torch.from_numpy(l) for l in batch["labels"]]) if c["mode"] == "predcls" else \
        dict(predict_logits=[torch.from_numpy(l) for l in batch["predict_logits"]])
    with torch.no_grad():
        ref = TP.predictor_forward(sd, boxes, pairs, x2d, d2d, c["mode"], **kw)
        ref_arg = ref[:, 1:].argmax(1)
        top2 = ref[:, 1:].topk(2, 1)[0]
        out = {}
        for m in modes:
            units, fn = make_mode(m)
            with LinearPatch(fn):
                y = TP.predictor_forward(sd, boxes, pairs, x2d, d2d, c["mode"], **kw)
            err = float((y - ref).abs().max() / ref.abs().max())
            flips = int((y[:, 1:].argmax(1) != ref_arg).sum())
            out[m] = dict(units=units, rel_err=err, argmax_flips=flips)
            print(f"{name:18s} {m:18s} units {units:3.1f}  rel err {err:9.2e}  argmax flips {flips:3d} / {len(ref)}", flush=True)
    return dict(pairs=int(len(ref)), min_top2_margin_rel=float(((top2[:, 0] - top2[:, 1]).min() / ref.abs().max())), modes=out)


def main():
    names = sys.argv[1:] or ["cfg1_predcls_vg", "sgdet_cap", "ragged_predcls"]
    res = {n: run_case(n, MODES) for n in names}
    path = os.path.join(ROOT, "profiles", "r2_precision_study.json")
    json.dump(res, open(path, "w"), indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
