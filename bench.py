#!/usr/bin/env python
"""bench.py — VETO relation-head throughput on B200 (the metric of BASELINE.json), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--precision bf16x3|bf16|fp32]

Headline workload: BASELINE.json configs[1] — VETO vanilla PredCls TRAINING step, IMS_PER_BATCH 12 per GPU, 20 GT
boxes / image (all 380 ordered pairs per image are under the 1024-pair cap of gtbox_relsample, so a step trains on
12 x 380 = 4560 pairs), VG 151/51, 592x800 images.  A "step" = gtbox_relsample (relation sampling) -> VETOFeatureExtractor (ROI gather)
-> VETOPredictor in train() mode with the reference's dropout rates -> rel_loss.backward() (incl. the ROIAlign
backward into the depth feature map) -> gradient all-reduce (N > 1, NCCL) -> clip_grad_norm 5.0 -> Adam step
(tools/relation_train_net.py:418-483).

  value : relation pairs / s trained, whole job, inputs resident in HBM, CUDA-event timed, max over ranks.
  e2e   : the same through the public API from pinned HOST buffers: H2D of the step's feature maps, boxes and labels
          and D2H of the loss inside the timed region.
  roofline : gemm_tc2_kernel (tcgen05, forward + both backward GEMMs of every encoder Linear): algorithmic FLOPs per
          launch / CUDA-event duration per launch (veto_profile_*), against MEASURED_PEAKS.json bf16_tflops_sustained.
  cpu_baseline : oracle/torch_port.py (the reference's formulation on torch CPU kernels + autograd) on one image.
  inference : configs[2] (SGDet-shaped inference, 32 images x 80 proposals = 202 240 pairs per step), same keys.

Multi-GPU: images are independent, so every rank runs its own batch (weak scaling); the only collective is the
gradient all-reduce of the training step.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "relation_pairs_per_sec"
UNIT = "pairs/s"
IMG_H, IMG_W = 592, 800
FLOP_PER_PAIR = 648.7e6          # forward, reference formulation (SURVEY.md §8d / BASELINE.md §3)
# configs[1]: training step
TR_IMAGES, TR_BOXES = 12, 20
TR_WORKLOAD = ("configs[1]: VETO vanilla PredCls training step, IMS_PER_BATCH 12 x 20 GT boxes (380 pairs/image, 4560 pairs/step), "
               "VG 151/51, 592x800 images, P2-P5 + depth NCHW fp32, dropout 0.1/0.35/0.35, Adam + clip 5.0")
# configs[2]: SGDet-shaped inference
N_IMAGES, N_BOXES, MAX_PAIRS = 32, 80, 8192
INF_WORKLOAD = ("configs[2]: SGDet-shaped inference, 32 images x 80 proposals (6320 pairs/image, MAX_PROPOSAL_PAIR 8192), "
                "VG 151/51, 592x800 images, P2-P5 + depth NCHW fp32; pairs -> ROI gather -> predictor -> post-processor")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default=os.environ.get("VETO_PRECISION", "bf16x3"), choices=["bf16x3", "bf16", "fp32"])
    ap.add_argument("--chunk", type=int, default=int(os.environ.get("VETO_CHUNK_PAIRS", "0")))
    ap.add_argument("--images", type=int, default=TR_IMAGES, help="training images per GPU per step")
    ap.add_argument("--inference-images", type=int, default=N_IMAGES)
    ap.add_argument("--cpu-sample-pairs", type=int, default=int(os.environ.get("VETO_CPU_SAMPLE_PAIRS", "2048")))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-inference", action="store_true", help="skip the configs[2] inference leg")
    ap.add_argument("--no-depth-backbone", action="store_true", help="skip the depth-backbone (SURVEY.md §8 f3) leg")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        return False

    def summary(self):
        sm = sorted(int(float(r[0])) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit())
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) >= 7 and r[3 + k].lower().startswith("active") for r in self.rows)]
        busy = [v for v in sm if v > 300] or sm
        return {"sm_mhz": busy[len(busy) // 2], "sm_max_mhz": int(float(self.rows[0][1])), "reasons": reasons,
                "samples": len(sm), "power_w_max": max(float(r[2]) for r in self.rows if len(r) >= 7)}


# ------------------------------------------------------------------------------------------------ CPU baselines
def cpu_train_leg(steps: int = 1, warmup: int = 0):
    """The reference's CPU training step for this workload, restated by oracle/torch_port.py (the reference's own
    formulation on torch CPU kernels, gradients by torch autograd — what the reference itself runs on CPU; the Python
    reference cannot travel to the GPU box): ONE image of the bench shape (20 boxes, 380 pairs): pair enumeration,
    ROI gather, predictor forward in train() mode, CE loss, backward.  Returns (pairs/s, cores, sample, s/step)."""
    import torch
    from oracle import torch_port as TP
    from veto_b200 import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    batch = synth.make_batch(1000, [TR_BOXES], H=IMG_H, W=IMG_W, mode="predcls")
    sd = TP.to_torch(synth.predictor_state(11))
    feats = [torch.from_numpy(f) for f in batch["feats"]]
    depth = torch.from_numpy(batch["depth"])
    boxes = [torch.from_numpy(b) for b in batch["boxes"]]
    labels = [torch.from_numpy(l) for l in batch["labels"]]
    R = TR_BOXES * (TR_BOXES - 1)
    import numpy as np
    from oracle import veto_oracle as O
    rel_mat = synth.make_relation_matrices(5, [TR_BOXES], 51, 10)[0]
    rng = np.random.default_rng(5)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        fg, fg_labels, bg, _ = O.gtbox_relsample_candidates(rel_mat)            # gtbox_relsample (sampling.py:54-107)
        bg = bg[rng.permutation(len(bg))][:1024 - len(fg)]
        pairs = [torch.from_numpy(np.concatenate([fg, bg]))]
        rel_labels = [torch.from_numpy(np.concatenate([fg_labels, np.zeros(len(bg), np.int64)]))]
        x2d, d2d = TP.pooler_forward(feats, depth, boxes)
        loss, grads, *_ = TP.train_step(sd, boxes, pairs, rel_labels, x2d, d2d, "predcls", labels=labels)
        dt = time.perf_counter() - t0
        assert bool(torch.isfinite(loss))
        if it >= warmup:
            times.append(dt)
    per_step = sorted(times)[len(times) // 2]
    sample = (f"1 image x {TR_BOXES} boxes = {R} pairs: relation sampling + ROI gather + VETOPredictor train() forward + CE loss "
              f"+ backward (no optimizer step); oracle/torch_port.py = the reference's formulation on torch CPU fp32 kernels "
              f"with torch autograd, {cores} threads")
    return R / per_step, cores, sample, per_step


def cpu_infer_leg(sample_pairs: int, steps: int = 1, warmup: int = 0):
    """configs[2] on the host cores: one image of 80 proposals, the predictor on the first `sample_pairs` of its 6320
    pairs (the reference materialises 262 KB per pair).  Returns (pairs/s, cores, sample, s/pair)."""
    import torch
    from oracle import torch_port as TP
    from veto_b200 import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    from oracle import veto_oracle as O
    batch = synth.make_batch(1000, [N_BOXES], H=IMG_H, W=IMG_W, mode="sgdet")
    synth.add_nms_fields(batch, 1300, relabel=False)
    sd = TP.to_torch(synth.predictor_state(11))
    feats = [torch.from_numpy(f) for f in batch["feats"]]
    depth = torch.from_numpy(batch["depth"])
    boxes = [torch.from_numpy(b) for b in batch["boxes"]]
    plog = [torch.from_numpy(l) for l in batch["predict_logits"]]
    times = []
    with torch.no_grad():
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            pairs = TP.prepare_test_pairs(batch["n_boxes"], MAX_PAIRS)
            x2d, d2d = TP.pooler_forward(feats, depth, boxes)
            t_fixed = time.perf_counter() - t0
            sub = [pairs[0][:sample_pairs]]
            t1 = time.perf_counter()
            logits = TP.predictor_forward(sd, boxes, sub, x2d, d2d, "sgdet", predict_logits=plog)
            O.postprocess_sgdet([logits.numpy()], batch["predict_logits"], [sub[0].numpy()], batch["boxes_per_cls"], 0.5)
            t_pairs = time.perf_counter() - t1
            assert bool(torch.isfinite(logits).all())
            if it >= warmup:
                times.append(t_fixed / len(pairs[0]) + t_pairs / len(sub[0]))
    per_pair = sorted(times)[len(times) // 2]
    sample = (f"1 image x {N_BOXES} proposals: pair enumeration + ROI gather (amortised over 6320 pairs) + predictor and "
              f"post-processor on the first {len(sub[0])} pairs; oracle/torch_port.py on torch CPU fp32 kernels, {cores} threads")
    return 1.0 / per_pair, cores, sample, per_pair


def cpu_depth_leg(steps: int = 1, warmup: int = 0):
    """The depth backbone's training forward + backward on the host cores: oracle/depth_port.py (the reference module IS
    torchvision's ResNet-18 trunk; restated on torch.nn.functional, gradients by autograd) on ONE 592x800 depth image.
    Returns (images/s, cores, sample)."""
    import numpy as np
    import torch
    from oracle import depth_port as DP
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = DP.synth_state(0)
    depth = DP.synth_depth(1, IMG_H, IMG_W)
    g = np.ones((1, 256) + DP.out_size(IMG_H, IMG_W), np.float32)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        DP.train_step(sd, depth, g)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    per = sorted(times)[len(times) // 2]
    return 1.0 / per, cores, (f"1 depth image 1x{IMG_H}x{IMG_W}: ResNetDepth train() forward + backward, oracle/depth_port.py on "
                              f"torch CPU fp32 kernels, {cores} threads")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.perf_counter()
    val, cores, sample, per_step = cpu_train_leg(steps=max(1, min(args.steps, 3)), warmup=min(args.warmup, 1))
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": TR_WORKLOAD, "l2": "host CPU run"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    if not args.no_inference:
        iv, _, isample, _ = cpu_infer_leg(args.cpu_sample_pairs)
        line["inference"] = {"value": iv, "unit": UNIT, "config": {"workload": INF_WORKLOAD}, "sample": isample}
    if not args.no_depth_backbone:
        dv, _, dsample = cpu_depth_leg(steps=2, warmup=1)
        line["depth_backbone"] = {"value": dv, "unit": "images/s", "sample": dsample}
    line["wall_s"] = time.perf_counter() - t0
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from tests import harness as H
    from veto_b200 import lib as L
    from veto_b200 import ops, registry, synth
    from veto_b200.distributed import allreduce_gradients, max_over_ranks
    from veto_b200.postprocess import make_roi_relation_post_processor
    from veto_b200.sampling import make_roi_relation_samp_processor
    from veto_b200.structures import BoxList

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L.require_device()
    rgb_hw, depth_hw = synth.fpn_shapes(IMG_H, IMG_W)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)" if peaks else "1.4 PFLOP/s sustained (of fallback)"
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    passes = {"bf16x3": 3, "bf16": 1, "fp32": 1}[args.precision]
    pin = lambda t: t.detach().cpu().pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = ops.launch_count()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1), dev) / steps, (ops.launch_count() - n0)

    def host_boxlists(bls):
        keys = ("labels", "predict_logits", "pred_scores", "pred_labels", "boxes_per_cls")
        return ([pin(b.bbox) for b in bls], [{k: pin(b.get_field(k)) for k in keys if b.has_field(k)} for b in bls])

    # =============================================================================== configs[1]: training step
    B = args.images
    R = B * TR_BOXES * (TR_BOXES - 1)
    batch = synth.make_batch(100 + rank, [TR_BOXES] * B, H=IMG_H, W=IMG_W, mode="predcls", features=False)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    feats_dev = [torch.randn((B, 256) + hw, generator=g, device=dev) for hw in rgb_hw]
    feats_dev.append(torch.zeros(B, 256, 1, 1, device=dev))           # P6: present in the reference's list, unused
    depth_dev = torch.relu(torch.randn((B, 256) + depth_hw, generator=g, device=dev)).requires_grad_(True)
    state = synth.predictor_state(11, spread=False)                   # torch-default-like init: a sane training start
    cfg = H.make_cfg(mode="predcls", precision=args.precision)
    pred = registry.make_roi_relation_predictor(cfg, 512)
    pred.load_state_dict(synth.to_torch_state(state), strict=True)
    pred = pred.to(dev).train()
    fe = registry.make_roi_box_feature_extractor(cfg, 256, for_relation=True).to(dev).train()
    samp = make_roi_relation_samp_processor(cfg)
    bls_dev = H.boxlists(batch, dev, 151)
    # ground-truth relation matrices (the targets' "relation" field): ~10 annotated relations per image, VG-like
    rel_mats_np = synth.make_relation_matrices(7 + rank, [TR_BOXES] * B, 51, 10)

    def make_targets(mats):
        out = []
        for bl, m in zip(bls_dev, mats):
            t = BoxList(bl.bbox, (IMG_W, IMG_H), "xyxy")
            t.add_field("relation", m)
            out.append(t)
        return out

    targets_dev = make_targets([torch.from_numpy(m).to(dev) for m in rel_mats_np])
    params = [p for p in pred.parameters() if p.requires_grad]
    opt = torch.optim.Adam(params, lr=1e-4 * B, fused=True)           # BASE_LR x IMS_PER_BATCH (relation_train_net.py:330-339)
    losses = []

    def train_step(feats, depth, bls, targets, pending=None):
        opt.zero_grad(set_to_none=True)
        depth.grad = None
        # relation sampling on the ground-truth boxes (relation_head.py:118-121 -> sampling.py:54-107): the annotated
        # pairs + randomly ordered background pairs; 20 boxes give 380 candidates, all under the 1024-pair cap.
        # `pending`: this step's sampling, enqueued one step earlier (it depends on the targets only), so that the
        # sampler's one host sync (row counts) never drains the GPU queue
        pend = pending if pending is not None else samp.gtbox_relsample_async(bls, targets)
        _, rel_labels, pairs, _ = pend.result()
        x2d, d2d, _, _ = fe(feats, bls, depth_features=depth)
        loss = pred(bls, pairs, rel_labels, None, roi_features=x2d, roi_depth_features=d2d)[2]["rel_loss"]
        loss.backward()
        allreduce_gradients(params)                                   # DDP's NCCL all-reduce (relation_train_net.py:372-380)
        torch.nn.utils.clip_grad_norm_(params, 5.0, foreach=True)     # GRAD_NORM_CLIP 5.0 (:475-481)
        opt.step()
        return loss.detach()

    resident = {"pending": None}

    def step_resident():
        pend = resident["pending"] or samp.gtbox_relsample_async(bls_dev, targets_dev)
        resident["pending"] = samp.gtbox_relsample_async(bls_dev, targets_dev)   # the NEXT step's sampling (every step samples anew)
        losses.append(train_step(feats_dev, depth_dev, bls_dev, targets_dev, pend))

    feats_host = [pin(f) for f in feats_dev]
    depth_host = pin(depth_dev)
    boxes_host, fields_host = host_boxlists(bls_dev)
    labels_host = [pin(t.get_field("relation")) for t in targets_dev]   # the targets' relation matrices
    h2d_bytes = sum(t.numel() * t.element_size() for t in feats_host + [depth_host] + boxes_host + labels_host)
    h2d_bytes += sum(t.numel() * t.element_size() for f in fields_host for t in f.values())
    loss_host = torch.empty(1, dtype=torch.float32).pin_memory()

    class InputPipe:
        """e2e input pipeline: two device-side input sets; a copy stream fills set k+1 from pinned host memory while
        the main stream works on set k (the data-loader prefetch every training / inference loop has).  Every
        acquire() issues exactly one H2D copy of a full step's inputs and hands out the set staged by the previous
        call (the first call stages its own)."""

        def __init__(self, make_slot, after_small=None):
            self.stream = torch.cuda.Stream(device=dev)
            self.after_small = after_small
            self.slots = []
            for _ in range(2):
                struct, copies = make_slot()
                self.slots.append({"struct": struct, "copies": copies, "ready": torch.cuda.Event(), "free": torch.cuda.Event(),
                                   "pending": None})
            self.next, self.staged = 0, None

        def _stage(self):
            sl = self.slots[self.next]
            self.stream.wait_event(sl["free"])                   # the step that used this set has finished
            with torch.cuda.stream(self.stream), torch.no_grad():
                order = sorted(sl["copies"], key=lambda dh: dh[1].numel() * dh[1].element_size())
                small = [dh for dh in order if dh[1].numel() * dh[1].element_size() <= (1 << 20)]
                for d, h in small:                               # boxes, labels, relation matrices first
                    d.copy_(h, non_blocking=True)
                if self.after_small is not None:                 # e.g. the relation sampling of this set, on this stream
                    sl["pending"] = self.after_small(sl)
                for d, h in order[len(small):]:
                    d.copy_(h, non_blocking=True)
                sl["ready"].record(self.stream)
            self.staged, self.next = self.next, self.next ^ 1

        def acquire(self):
            if self.staged is None:
                self._stage()
            sl = self.slots[self.staged]
            self._stage()                                        # this call's H2D copy: the NEXT step's inputs
            torch.cuda.current_stream().wait_event(sl["ready"])
            return sl

        @staticmethod
        def release(sl):
            sl["free"].record()

    def dev_like(h):
        return torch.empty(h.shape, dtype=h.dtype, device=dev)

    def boxlist_slot(boxes_h, fields_h):
        bls, copies = [], []
        for bb, ff in zip(boxes_h, fields_h):
            bl = BoxList(dev_like(bb), (IMG_W, IMG_H), "xyxy")
            copies.append((bl.bbox, bb))
            for k, v in ff.items():
                d = dev_like(v)
                bl.add_field(k, d)
                copies.append((d, v))
            bls.append(bl)
        return bls, copies

    def train_slot():
        feats = [dev_like(f) for f in feats_host]
        depth = dev_like(depth_host).requires_grad_(True)
        bls, copies = boxlist_slot(boxes_host, fields_host)
        mats = [dev_like(l) for l in labels_host]
        targets = []
        for bl, m in zip(bls, mats):
            t = BoxList(bl.bbox, (IMG_W, IMG_H), "xyxy")
            t.add_field("relation", m)
            targets.append(t)
        copies += list(zip(feats, feats_host)) + [(depth, depth_host)] + list(zip(mats, labels_host))
        return (feats, depth, bls, targets), copies

    def sample_with_inputs(sl):
        """Relation sampling of an input set as part of the input pipeline: on the copy stream, right behind the H2D
        copy of the set's boxes / relation matrices and ahead of its feature maps."""
        pend = samp.gtbox_relsample_async(sl["struct"][2], sl["struct"][3])
        for t in [pend.pairs, pend.labels] + list(pend.binaries):
            t.record_stream(torch.cuda.default_stream(dev))      # consumed by the training step on the main stream
        return pend

    train_pipe = InputPipe(train_slot, after_small=None if os.environ.get("VETO_BENCH_SYNC_SAMPLER") else sample_with_inputs)

    def step_e2e():
        sl = train_pipe.acquire()
        feats, depth, bls, targets = sl["struct"]
        pend, sl["pending"] = sl["pending"], None
        loss = train_step(feats, depth, bls, targets, pend)
        train_pipe.release(sl)
        loss_host.copy_(loss.reshape(1), non_blocking=True)

    with ClockSampler(local) as clk:
        ms_step, launches = timed(step_resident, args.steps, args.warmup)
    clocks = clk.summary()
    value = world * R / ms_step * 1e3
    loss_first, loss_last = float(losses[0]), float(losses[-1])
    ms_e2e, _ = timed(step_e2e, max(1, args.steps), 2)
    e2e_value = world * R / ms_e2e * 1e3

    # ---- per-stage device time of one step (CUDA events around every launch) -> roofline of the GEMM kernel
    torch.cuda.synchronize()
    with ops.StageTimer() as st:
        step_resident()
    M = R * 19
    enc_flops = 6 * 2.0 * M * (1728 * 576 + 576 * 576 + 2 * 576 * 1152)   # forward encoder Linears, 6 full layers
    fwd_tags = ["gemm_qkv", "gemm_out", "gemm_ff1", "gemm_ff2"]
    g_ms = sum(st.ms.get(k, 0.0) for k in fwd_tags + ["bwd_dgrad", "bwd_wgrad"])
    g_launch = sum(st.launches.get(k, 0) for k in fwd_tags + ["bwd_dgrad", "bwd_wgrad"])
    algo_flops = 3.0 * enc_flops                                           # forward + input-gradient + weight-gradient
    achieved = algo_flops / (g_ms * 1e-3) / 1e12 if g_ms > 0 else 0.0
    total_stage_ms = sum(st.ms.values())
    traffic, traffic_src = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "gemm_traffic.json")))
        traffic, traffic_src = tj.get("train_" + args.precision), tj.get("train_source")
    except Exception:
        pass
    roofline = {
        "bound": "tensor",
        "kernel": "gemm_tc2_kernel / gemm_tn2_kernel (tcgen05.mma cta_group::2, TMEM accumulators, TMA): forward + dX, and dW GEMMs" if args.precision != "fp32" else "gemm_simt_kernel",
        "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf, "traffic": traffic,
        "peak_source": peak_src, "traffic_source": traffic_src,
        "algorithmic_flops_per_launch": algo_flops / max(g_launch, 1), "launches_per_step": g_launch,
        "avg_launch_ms": g_ms / max(g_launch, 1), "share_of_step": g_ms / total_stage_ms if total_stage_ms else None,
        "executed_mma_tflops": achieved * passes, "frac_executed": achieved * passes / peak_tf,
        "note": ("algorithmic = 2*M*N*K of the reference's fp32 encoder Linears x 3 (forward, dX, dW); bf16x3 executes 3 bf16 MMAs "
                 "per product; bwd_dgrad / bwd_wgrad also hold the split-K reductions, the patch-projection and the small fp32 GEMMs"),
        "stage_ms": {k: round(v, 3) for k, v in st.ms.items()},
        "stage_launches": dict(st.launches),
    }

    # HBM-bound kernels of the training step: algorithmic bytes (every operand read / written once) / stage time
    nb_tr = B * TR_BOXES
    tr_bytes = {
        "roi_gather": nb_tr * 131072,
        "layernorm": 12 * M * 576 * (4 + 4),                               # fp32 in, bf16 hi + lo out
        "bwd_layernorm": 12 * M * 576 * (3 * 4 + 4 + 4),                   # x, dy, residual in; dx fp32 + operand hi/lo out
        "attention": 6 * M * (1728 * 4 + 576 * 4),                         # qkv fp32 in, output hi + lo
        "bwd_attention": 6 * M * (1728 * 4 + 576 * 4 + 1728 * 4),          # qkv, dO in; d_qkv hi + lo out
        "tokens": R * 19 * 576 * 4,
    }
    train_hbm = {k: {"ms": round(st.ms[k], 3), "algorithmic_bytes": int(v), "achieved_gbs": round(v / st.ms[k] / 1e6, 1),
                     "frac_of_measured_hbm": round(v / st.ms[k] / 1e6 / hbm_peak, 3)}
                 for k, v in tr_bytes.items() if st.ms.get(k)}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, cores, sample, _ = cpu_train_leg()
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}

    # release the training state before the inference leg
    train_ws_gb = sum(t.numel() for t in ops._workspaces.values()) / 1e9
    resident["pending"] = None
    del opt, params, pred, fe, feats_dev, depth_dev, feats_host, depth_host, train_pipe
    ops._workspaces.clear()
    torch.cuda.empty_cache()

    # =============================================================================== f3: the depth backbone
    depth_leg = None
    if not args.no_depth_backbone:
        from oracle import depth_port as DP          # synthetic state / image generators only (numpy)
        from veto_b200 import depth_backbone as DB
        dmodel = DB.build_resnet18_depth(cfg).to(dev).train()
        dmodel.load_state_dict({k: torch.from_numpy(v) for k, v in DP.synth_state(0).items()})
        dimg = torch.from_numpy(DP.synth_depth(B, IMG_H, IMG_W, seed=rank)).to(dev)
        dgrad = torch.randn((B, 256) + ops.depth_backbone_out_size(IMG_H, IMG_W), generator=g, device=dev)

        def depth_fwd():
            with torch.no_grad():
                dmodel(dimg)

        def depth_step():
            dmodel(dimg).backward(dgrad)

        d_fwd_ms, _ = timed(depth_fwd, max(3, args.steps // 2), 2)
        d_ms, _ = timed(depth_step, max(3, args.steps // 2), 2)
        ws_gb = L.load().veto_depth_backbone_workspace_bytes(L.PRECISIONS[args.precision], B, IMG_H, IMG_W, 1) / 1e9
        dcpu = None
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            v, cores, sample = cpu_depth_leg(steps=2, warmup=1)
            dcpu = {"value": v, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample}
        # convolution FLOPs of one image (forward; the backward's input + weight gradients are twice that, conv1 has no
        # input gradient)
        ho, wo = (IMG_H + 6 - 7) // 2 + 1, (IMG_W + 6 - 7) // 2 + 1       # conv1 7x7 / 2, pad 3, one input channel
        conv1_flop = 2.0 * ho * wo * 64 * 49
        conv_flop = conv1_flop
        ph, pw = (ho - 1) // 2 + 1, (wo - 1) // 2 + 1                       # max-pool 3x3 / 2
        # layer1 keeps the pooled size, layer2 / layer3 halve it; two blocks of two 3x3 each + one 1x1 where the layer strides
        for cin, cout, div in ((64, 64, 1), (64, 128, 2), (128, 256, 4)):
            oh, ow = (ph - 1) // div + 1, (pw - 1) // div + 1
            first = 2.0 * oh * ow * cout * cin * 9 + (2.0 * oh * ow * cout * cin if cin != cout else 0.0)
            conv_flop += first + 3 * 2.0 * oh * ow * cout * cout * 9
        step_flop = B * (3 * conv_flop - conv1_flop)
        d_tf = step_flop / d_ms / 1e9
        depth_leg = {"metric": "depth_images_per_sec", "value": world * B / d_ms * 1e3, "unit": "images/s",
                     "ms_fwd_bwd": d_ms, "ms_fwd": d_fwd_ms,
                     "roofline": {"bound": "tensor", "achieved": d_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": d_tf / peak_tf,
                                  "executed_mma_tflops": d_tf * passes, "frac_executed": d_tf * passes / peak_tf,
                                  "note": "convolution FLOPs (forward + input + weight gradients) / step time; the step is "
                                          "bound by operand delivery and the BatchNorm HBM passes, DESIGN.md 7a"},
                     "config": {"workload": f"SURVEY.md §8 f3: ResNetDepth (R-18-C4) train() forward + backward, {B} depth images "
                                            f"1x{IMG_H}x{IMG_W} per GPU -> [B,256,H/16,W/16]", "precision": args.precision},
                     "workspace_gb": round(ws_gb, 2), "cpu_baseline": dcpu}
        del dmodel, dimg, dgrad
        torch.cuda.empty_cache()

    # =============================================================================== configs[2]: inference
    inference = None
    if not args.no_inference:
        Bi = args.inference_images
        Ri = Bi * N_BOXES * (N_BOXES - 1)
        ibatch = synth.make_batch(200 + rank, [N_BOXES] * Bi, H=IMG_H, W=IMG_W, mode="sgdet", features=False)
        synth.add_nms_fields(ibatch, 300 + rank, relabel=False)          # the detector's per-class boxes (late NMS input)
        ifeats = [torch.randn((Bi, 256) + hw, generator=g, device=dev) for hw in rgb_hw]
        idepth = torch.relu(torch.randn((Bi, 256) + depth_hw, generator=g, device=dev))
        icfg = H.make_cfg(mode="sgdet", max_pairs=MAX_PAIRS, precision=args.precision, chunk_pairs=args.chunk)
        ipred = H.build_predictor(icfg, synth.predictor_state(11), dev)
        ife = registry.make_roi_box_feature_extractor(icfg, 256, for_relation=True).to(dev).eval()
        isamp = make_roi_relation_samp_processor(icfg)
        ipost = make_roi_relation_post_processor(icfg)
        ibls = H.boxlists(ibatch, dev, 151)

        def infer_head(feats, depth, bls):
            """ROIRelationHead.forward at test time (relation_head.py:134-243): candidate pairs -> ROI gather ->
            predictor -> post-processor (late per-class NMS of the objects, triple scores, per-image ranking)."""
            pairs = isamp.prepare_test_pairs(dev, bls)
            x2d, d2d, _, _ = ife(feats, bls, depth_features=depth)
            rel = ipred(bls, pairs, None, None, roi_features=x2d, roi_depth_features=d2d)[1]
            return ipost((rel, [b.get_field("predict_logits") for b in bls]), pairs, bls)

        def infer_resident():
            with torch.no_grad():
                return infer_head(ifeats, idepth, ibls)

        ifeats_host = [pin(f) for f in ifeats]
        idepth_host = pin(idepth)
        iboxes_host, ifields_host = host_boxlists(ibls)
        ih2d = sum(t.numel() * t.element_size() for t in ifeats_host + [idepth_host] + iboxes_host)
        ih2d += sum(t.numel() * t.element_size() for f in ifields_host for t in f.values())
        logits_host = torch.empty((Ri, 51), dtype=torch.float32).pin_memory()      # ranked class probabilities
        pairs_host = torch.empty((Ri, 2), dtype=torch.int64).pin_memory()           # ranked pairs

        def infer_slot():
            feats = [dev_like(f) for f in ifeats_host]
            depth = dev_like(idepth_host)
            bls, copies = boxlist_slot(iboxes_host, ifields_host)
            copies += list(zip(feats, ifeats_host)) + [(depth, idepth_host)]
            return (feats, depth, bls), copies

        infer_pipe = InputPipe(infer_slot)

        def infer_e2e():
            with torch.no_grad():
                sl = infer_pipe.acquire()
                feats, depth, bls = sl["struct"]
                res = infer_head(feats, depth, bls)
                infer_pipe.release(sl)
                logits_host.copy_(torch.cat([r.get_field("pred_rel_scores") for r in res]), non_blocking=True)
                pairs_host.copy_(torch.cat([r.get_field("rel_pair_idxs") for r in res]), non_blocking=True)

        isteps = max(2, min(args.steps, 5))
        ims, ilaunches = timed(infer_resident, isteps, max(3, min(args.warmup, 3)))
        ims_e2e, _ = timed(infer_e2e, max(1, min(args.steps, 3)), 1)
        torch.cuda.synchronize()
        with ops.StageTimer() as ist:
            infer_resident()
        Mi = Ri * 19
        iflops = (5 * 2.0 * Mi * (1728 * 576 + 576 * 576 + 2 * 576 * 1152)           # five full layers
                  + 2.0 * Mi * 1152 * 576 + 2.0 * Ri * (576 * 576 * 2 + 2 * 576 * 1152))  # last layer: K, V of all rows; CLS row only elsewhere
        ig_ms = sum(ist.ms.get(k, 0.0) for k in fwd_tags)
        ig_launch = sum(ist.launches.get(k, 0) for k in fwd_tags)
        iach = iflops / (ig_ms * 1e-3) / 1e12 if ig_ms > 0 else 0.0
        nb_total = Bi * N_BOXES
        rows_full = 5 * Mi
        hbm_bytes = {
            "pairs": Ri * (16 + 16 + 8),
            "roi_gather": nb_total * 131072,
            "tokens": Ri * 19 * 576 * 4,
            "layernorm": (2 * rows_full + Mi + Ri) * 4608,
            "attention": rows_full * (6912 + 2304) + Mi * 4608 + Ri * 4608,
        }
        hbm_kernels = {k: {"ms": round(ist.ms[k], 3), "algorithmic_bytes": int(v), "achieved_gbs": round(v / ist.ms[k] / 1e6, 1),
                           "frac_of_measured_hbm": round(v / ist.ms[k] / 1e6 / hbm_peak, 3)}
                       for k, v in hbm_bytes.items() if ist.ms.get(k)}
        icpu = None
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            v, cores, sample, _ = cpu_infer_leg(args.cpu_sample_pairs)
            icpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
        inference = {
            "metric": METRIC, "value": world * Ri / ims * 1e3, "unit": UNIT, "ms_per_step": ims, "steps": isteps,
            "images_per_sec": world * Bi / ims * 1e3,
            "config": {"workload": INF_WORKLOAD, "images_per_gpu": Bi, "pairs_per_step_per_gpu": Ri,
                       "chunk_pairs": ipred.chunk_pairs or "library default"},
            "e2e": {"value": world * Ri / ims_e2e * 1e3, "unit": UNIT, "h2d_bytes_per_step": ih2d,
                    "d2h_bytes_per_step": logits_host.numel() * 4 + pairs_host.numel() * 8, "ms_per_step": ims_e2e,
                    "pipeline": "double-buffered H2D on a copy stream, overlapped with the previous step"},
            "gpu_launches": ilaunches,
            "tflops_reference_formulation": world * Ri / ims * 1e3 * FLOP_PER_PAIR / 1e12,
            "roofline": {"bound": "tensor", "achieved": iach, "peak": peak_tf, "unit": "TFLOP/s", "frac": iach / peak_tf,
                         "executed_mma_tflops": iach * passes, "frac_executed": iach * passes / peak_tf,
                         "launches_per_step": ig_launch, "stage_ms": {k: round(v, 3) for k, v in ist.ms.items()}},
            "hbm_kernels": hbm_kernels, "cpu_baseline": icpu,
        }

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"bf16x3": "bf16x3 (split-bf16 tensor-core products, fp32 accumulate; fp32-grade)", "bf16": "bf16",
                      "fp32": "f32"}[args.precision],
            "data": "synthetic",
            "config": {"workload": TR_WORKLOAD, "images_per_gpu": B, "pairs_per_step_per_gpu": R, "precision": args.precision,
                       "parallelism": f"image-sharded data parallel x{world}, one flat NCCL gradient all-reduce per step",
                       "optimizer": "Adam (torch fused), clip_grad_norm 5.0",
                       "l2": "inputs larger than L2 (0.52 GB of feature maps per step; 14 GB of saved activations)"},
            "images_per_sec": world * B / ms_step * 1e3,
            "tflops_reference_formulation": value * 3 * FLOP_PER_PAIR / 1e12,
            "loss_first_last": [loss_first, loss_last], "train_workspace_gb": round(train_ws_gb, 2),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e,
                    "pipeline": "double-buffered: the H2D copy of step k+1 (copy stream) overlaps the training of step k"},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "hbm_peak_gbs": hbm_peak,
            "hbm_kernels": train_hbm,
            "cpu_baseline": cpu, "inference": inference, "depth_backbone": depth_leg,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
