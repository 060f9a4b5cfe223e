#!/usr/bin/env python
"""Inference probe for GPU experiments (not the bench): configs[2]-shaped SGDet inference on one B200.

    python tools/infer_probe.py [--images 32] [--boxes 80] [--chunks 0,997,3988] [--precision bf16x3] [--steps 3]
                                [--once]   (one un-timed step only: the command to put under ncu)

Prints one JSON line per chunk size: ms / step (CUDA events), pairs / s, and the per-stage device times of one step.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=32)
    ap.add_argument("--boxes", type=int, default=80)
    ap.add_argument("--max-pairs", type=int, default=8192)
    ap.add_argument("--chunks", default="0")
    ap.add_argument("--precision", default="bf16x3")
    ap.add_argument("--mode", default="sgdet")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--once", action="store_true")
    args = ap.parse_args()

    import torch
    from veto_b200 import lib as L
    from veto_b200 import ops, registry, synth, workloads as WL
    from veto_b200.postprocess import make_roi_relation_post_processor
    from veto_b200.sampling import make_roi_relation_samp_processor

    L.require_device()
    dev = torch.device("cuda", 0)
    H, W = 592, 800
    B, N = args.images, args.boxes
    batch = synth.make_batch(200, [N] * B, H=H, W=W, mode=args.mode, features=False)
    if args.mode == "sgdet":
        synth.add_nms_fields(batch, 300, relabel=False)
    g = torch.Generator(device=dev).manual_seed(1234)
    feats, depth = WL.random_features(B, H, W, dev, g)
    bls = WL.boxlists(batch, dev, 151)
    state = synth.predictor_state(11)
    for chunk in [int(c) for c in args.chunks.split(",")]:
        cfg = WL.make_cfg(mode=args.mode, max_pairs=args.max_pairs, precision=args.precision, chunk_pairs=chunk)
        pred = WL.build_predictor(cfg, state, dev)
        fe = registry.make_roi_box_feature_extractor(cfg, 256, for_relation=True).to(dev).eval()
        samp = make_roi_relation_samp_processor(cfg)
        post = make_roi_relation_post_processor(cfg)

        def step():
            with torch.no_grad():
                pairs = samp.prepare_test_pairs(dev, bls)
                x2d, d2d, _, _ = fe(feats, bls, depth_features=depth)
                rel = pred(bls, pairs, None, None, roi_features=x2d, roi_depth_features=d2d)[1]
                return post((rel, [b.get_field("predict_logits") for b in bls]), pairs, bls), pairs

        if args.once:
            step()
            torch.cuda.synchronize()
            continue
        for _ in range(args.warmup):
            _, pairs = step()
        R = sum(len(p) for p in pairs)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        with ops.StageTimer() as st:
            step()
        print(json.dumps({"chunk": chunk or "default", "images": B, "pairs": R, "ms_per_step": round(ms, 2),
                          "pairs_per_s": round(R / ms * 1e3), "precision": args.precision,
                          "stage_ms": {k: round(v, 2) for k, v in st.ms.items()}}), flush=True)
        del pred
        ops._workspaces.clear()
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
