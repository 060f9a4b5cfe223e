#!/usr/bin/env python
"""bench.py — VETO relation-head throughput on B200 (the metric of BASELINE.json), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--precision f16c8|bf16x3|f16|bf16|fp32]

Headline workload (top-level keys): BASELINE.json configs[2], the largest single-GPU configuration — SGDet-shaped
INFERENCE, 32 images x 80 proposals = 202 240 candidate pairs per step per GPU (MAX_PROPOSAL_PAIR 8192), VG 151/51,
592x800 images.  A "step" = ROIRelationHead.forward at test time (relation_head.py:134-243): prepare_test_pairs -> ROI
gather of the RGB + depth features -> VETOPredictor -> PostProcessor (late per-class NMS, triple scores, per-image ranking).

  value : relation pairs / s, whole job, inputs resident in HBM, CUDA-event timed, max over ranks.
  e2e   : the same through the public API from pinned HOST buffers: H2D of the step's feature maps / boxes / detector
          fields and D2H of the ranked pairs + class probabilities inside the timed region (double-buffered copy stream).
  roofline : the dominant kernel, gemm_tc2_kernel (tcgen05 encoder GEMMs): FLOPs of the Linears it executed / its
          CUDA-event time (veto_profile_*), against MEASURED_PEAKS.json bf16_tflops_sustained.
  cpu_baseline : the reference's formulation on the host cores (oracle/torch_port.py), bounded sample of the same
          workload: 3 warm-ups + >= 5 timed repetitions, median.

Sub-objects, each with its own value / e2e / roofline or hbm figures / cpu_baseline:
  train     configs[1]: PredCls training step, 12 images x 20 GT boxes = 4560 pairs; the step runs depth image -> depth
            backbone (R-18-C4, trained) -> relation sampling -> ROI gather -> VETOPredictor train() -> backward through
            the head, the ROIAlign and the backbone -> gradient all-reduce (N > 1) -> clip -> Adam over both modules.
  meet_gqa  configs[3]: VETOPredictor_MEET PredCls inference, GQA 201/101, 16 images x 20 boxes = 6080 pairs, 4 group heads.
  sweep96   configs[4]: 96 mixed images (48 PredCls x 20 boxes + 48 SGDet x 80 proposals), sharded by pair count over the
            ranks (strong scaling: the 96 images are the whole job at every N).

Multi-GPU: images are independent, so the headline gives every rank its own batch (weak scaling) and needs no
collective; the training step's only collective is the in-place gradient all-reduce.

--impl reference: the reference's CPU path (restated on torch CPU kernels, oracle/torch_port.py) on the SAME headline
workload — per image, on a bounded sample (one 80-proposal image, the predictor on its first --cpu-sample-pairs pairs) —
K timed steps after W warm-ups, as the flags say.
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
ENC_MAC_PER_ROW = 1728 * 576 + 576 * 576 + 2 * 576 * 1152     # encoder Linears of one token row of one layer
# configs[2]: SGDet-shaped inference (headline)
N_IMAGES, N_BOXES, MAX_PAIRS = 32, 80, 8192
INF_WORKLOAD = ("configs[2]: SGDet-shaped inference, 32 images x 80 proposals (6320 pairs/image = 202 240 pairs/step, "
                "MAX_PROPOSAL_PAIR 8192), VG 151/51, 592x800 images, P2-P5 + depth NCHW fp32; prepare_test_pairs -> ROI gather -> "
                "VETOPredictor -> PostProcessor")
# configs[1]: training step
TR_IMAGES, TR_BOXES = 12, 20
TR_WORKLOAD = ("configs[1]: VETO vanilla PredCls training step, IMS_PER_BATCH 12 x 20 GT boxes (380 pairs/image, 4560 pairs/step), "
               "VG 151/51, 592x800 images; depth image -> depth backbone (trained) -> gtbox_relsample -> ROI gather -> predictor "
               "train() (dropout 0.1/0.35/0.35) -> backward through head + ROIAlign + backbone -> all-reduce -> clip 5.0 -> Adam")
# configs[3]: MEET GQA
MEET_IMAGES, MEET_BOXES = 16, 20
MEET_WORKLOAD = ("configs[3]: VETOPredictor_MEET PredCls inference, GQA 201 obj / 101 predicates, divide4 group heads "
                 "(7+12+22+67 outputs), 16 images x 20 boxes = 6080 pairs/step; pairs -> ROI gather -> predictor -> MEET post-processor")
# configs[4]: 96-image sweep
SWEEP_WORKLOAD = ("configs[4]: 96 mixed images = 48 PredCls x 20 GT boxes + 48 SGDet x 80 proposals (321 600 pairs), sharded by "
                  "pair count (LPT) over the ranks; every image through pairs -> ROI gather -> predictor -> post-processor")
HEADLINE_CONFIG = {"workload": INF_WORKLOAD, "images_per_gpu": N_IMAGES, "proposals_per_image": N_BOXES,
                   "pairs_per_step_per_gpu": N_IMAGES * N_BOXES * (N_BOXES - 1),
                   "l2": "inputs larger than L2: 1.39 GB of feature maps per step, > 2 GB of per-chunk activations"}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default=os.environ.get("VETO_PRECISION", "f16c8"),
                    choices=["f16c8", "bf16x3", "f16", "bf16", "fp32"])
    ap.add_argument("--chunk", type=int, default=int(os.environ.get("VETO_CHUNK_PAIRS", "0")))
    ap.add_argument("--cpu-sample-pairs", type=int, default=int(os.environ.get("VETO_CPU_SAMPLE_PAIRS", "1024")))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--legs", default="train,meet_gqa,sweep96", help="comma list of sub-legs to run next to the headline")
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


# ------------------------------------------------------------------------------------------------ CPU legs (oracle/)
def _median(xs):
    xs = sorted(xs)
    return xs[len(xs) // 2]


def cpu_infer_leg(sample_pairs: int, steps: int, warmup: int, n_boxes: int = N_BOXES, mode: str = "sgdet"):
    """One image of the headline workload on the host cores: pair enumeration + ROI gather of the whole image (amortised
    over all of its pairs), predictor + post-processor on the first `sample_pairs` pairs (the reference materialises
    262 KB per pair).  Returns (pairs/s, cores, sample description, seconds per pair, repetitions run)."""
    import torch
    from oracle import torch_port as TP
    from oracle import veto_oracle as O
    from veto_b200 import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    batch = synth.make_batch(1000, [n_boxes], H=IMG_H, W=IMG_W, mode=mode)
    if mode == "sgdet":
        synth.add_nms_fields(batch, 1300, relabel=False)
    sd = TP.to_torch(synth.predictor_state(11))
    feats = [torch.from_numpy(f) for f in batch["feats"]]
    depth = torch.from_numpy(batch["depth"])
    boxes = [torch.from_numpy(b) for b in batch["boxes"]]
    kw = (dict(predict_logits=[torch.from_numpy(l) for l in batch["predict_logits"]]) if mode == "sgdet"
          else dict(labels=[torch.from_numpy(l) for l in batch["labels"]]))
    times = []
    with torch.no_grad():
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            pairs = TP.prepare_test_pairs(batch["n_boxes"], MAX_PAIRS)
            x2d, d2d = TP.pooler_forward(feats, depth, boxes)
            t_fixed = time.perf_counter() - t0
            sub = [pairs[0][:sample_pairs]]
            t1 = time.perf_counter()
            logits = TP.predictor_forward(sd, boxes, sub, x2d, d2d, mode, **kw)
            if mode == "sgdet":
                O.postprocess_sgdet([logits.numpy()], batch["predict_logits"], [sub[0].numpy()], batch["boxes_per_cls"], 0.5)
            else:
                ol = torch.full((n_boxes, 151), -1000.0)
                ol[torch.arange(n_boxes), torch.from_numpy(batch["labels"][0])] = 1000.0
                O.postprocess([logits.numpy()], [ol.numpy()], [sub[0].numpy()])
            t_pairs = time.perf_counter() - t1
            assert bool(torch.isfinite(logits).all())
            if it >= warmup:
                times.append(t_fixed / len(pairs[0]) + t_pairs / len(sub[0]))
    per_pair = _median(times)
    sample = (f"1 image x {n_boxes} {'proposals' if mode == 'sgdet' else 'GT boxes'} ({len(pairs[0])} pairs): pair enumeration + ROI "
              f"gather (amortised over the image's pairs) + predictor and post-processor on the first {len(sub[0])} pairs; "
              f"oracle/torch_port.py = the reference's formulation on torch CPU fp32 kernels, {cores} threads; "
              f"{warmup} warm-ups + {steps} timed, median")
    return 1.0 / per_pair, cores, sample, per_pair, len(times)


def cpu_train_leg(steps: int, warmup: int):
    """The reference's CPU training step restated by oracle/torch_port.py + oracle/depth_port.py, ONE image of the bench
    shape (20 boxes, 380 pairs): depth backbone forward, relation sampling, ROI gather, predictor train() forward, CE
    loss, backward through head + backbone.  Returns (pairs/s, cores, sample, s/step)."""
    import numpy as np
    import torch
    from oracle import depth_port as DP
    from oracle import torch_port as TP
    from oracle import veto_oracle as O
    from veto_b200 import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    batch = synth.make_batch(1000, [TR_BOXES], H=IMG_H, W=IMG_W, mode="predcls")
    sd = TP.to_torch(synth.predictor_state(11))
    feats = [torch.from_numpy(f) for f in batch["feats"]]
    boxes = [torch.from_numpy(b) for b in batch["boxes"]]
    labels = [torch.from_numpy(l) for l in batch["labels"]]
    dstate = {k: torch.from_numpy(np.asarray(v)).clone().requires_grad_(v.dtype == np.float32 and "running" not in k)
              for k, v in synth.depth_state(0).items()}
    dimg = torch.from_numpy(synth.depth_images(1, IMG_H, IMG_W))
    R = TR_BOXES * (TR_BOXES - 1)
    rel_mat = synth.make_relation_matrices(5, [TR_BOXES], 51, 10)[0]
    rng = np.random.default_rng(5)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        depth = DP.forward(dstate, dimg, True)
        fg, fg_labels, bg, _ = O.gtbox_relsample_candidates(rel_mat)            # gtbox_relsample (sampling.py:54-107)
        bg = bg[rng.permutation(len(bg))][:1024 - len(fg)]
        pairs = [torch.from_numpy(np.concatenate([fg, bg]))]
        rel_labels = [torch.from_numpy(np.concatenate([fg_labels, np.zeros(len(bg), np.int64)]))]
        x2d, d2d = TP.pooler_forward(feats, depth.detach(), boxes)
        loss, grads, g_d2d, *_ = TP.train_step(sd, boxes, pairs, rel_labels, x2d, d2d, "predcls", labels=labels)
        # the backbone's backward (the ROIAlign backward in between is ~1 % of the step and left out of this sample)
        depth.backward(torch.ones_like(depth))
        for v in dstate.values():
            v.grad = None
        dt = time.perf_counter() - t0
        assert bool(torch.isfinite(loss))
        if it >= warmup:
            times.append(dt)
    per_step = _median(times)
    sample = (f"1 image x {TR_BOXES} boxes = {R} pairs: depth backbone forward + backward, relation sampling, ROI gather, "
              f"VETOPredictor train() forward + CE loss + backward (no optimizer step); oracle/torch_port.py + depth_port.py on "
              f"torch CPU fp32 kernels with torch autograd, {cores} threads; {warmup} warm-ups + {steps} timed, median")
    return R / per_step, cores, sample, per_step


def cpu_meet_leg(steps: int, warmup: int):
    """configs[3] on the host cores: one GQA image of 20 boxes (380 pairs) through the MEET trunk + the 4 group heads +
    the 'ensemble' merge of the post-processor (oracle restatements)."""
    import torch
    from oracle import torch_port as TP
    from oracle import veto_oracle as O
    from veto_b200 import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sizes = synth.GROUP_SPLITS[("GQA", "divide4")]
    batch = synth.make_batch(1000, [MEET_BOXES], H=IMG_H, W=IMG_W, num_obj=201, mode="predcls")
    sd_np = synth.meet_state(14, 201, sizes)
    sd = TP.to_torch({k[len("model."):]: v for k, v in sd_np.items() if k.startswith("model.")})
    sd["rel_out.weight"] = torch.cat([sd[f"rel_out.{k}.weight"] for k in range(len(sizes))], 0)
    sd["rel_out.bias"] = torch.cat([sd[f"rel_out.{k}.bias"] for k in range(len(sizes))], 0)
    feats = [torch.from_numpy(f) for f in batch["feats"]]
    depth = torch.from_numpy(batch["depth"])
    boxes = [torch.from_numpy(b) for b in batch["boxes"]]
    labels = [torch.from_numpy(l) for l in batch["labels"]]
    incre = O.incre_idx_list(sizes, 101)
    times = []
    with torch.no_grad():
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            pairs = TP.prepare_test_pairs(batch["n_boxes"], 2048)
            x2d, d2d = TP.pooler_forward(feats, depth, boxes)
            logits = TP.predictor_forward(sd, boxes, pairs, x2d, d2d, "predcls", labels=labels).numpy()
            gl, off = {}, 0
            for k, n in enumerate(sizes):
                gl["group_%d" % k] = logits[:, off:off + n + 2]
                off += n + 2
            ol = torch.full((MEET_BOXES, 201), -1000.0)
            ol[torch.arange(MEET_BOXES), labels[0]] = 1000.0
            O.postprocess_meet(gl, ol.numpy(), pairs[0].numpy(), incre)
            if it >= warmup:
                times.append(time.perf_counter() - t0)
    per = _median(times)
    R = MEET_BOXES * (MEET_BOXES - 1)
    return R / per, cores, (f"1 GQA image x {MEET_BOXES} boxes = {R} pairs: pairs + ROI gather + MEET trunk + 4 group heads + "
                            f"'ensemble' post-processor; oracle restatements on torch CPU fp32 kernels, {cores} threads; "
                            f"{warmup} warm-ups + {steps} timed, median")


def run_reference(args):
    """The reference arm: the SAME headline workload (configs[2]) on the host cores, per image on a bounded sample;
    exactly --steps timed repetitions after --warmup warm-ups."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.perf_counter()
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    val, cores, sample, per_pair, ran = cpu_infer_leg(args.cpu_sample_pairs, steps, warmup)
    pairs_step = N_IMAGES * N_BOXES * (N_BOXES - 1)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": ran,
            "warmup": warmup, "ms_per_step": per_pair * pairs_step * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": dict(HEADLINE_CONFIG),
            "images_per_sec": val / (N_BOXES * (N_BOXES - 1)),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "ms_per_step = the per-pair time of the sample x the 202 240 pairs of one headline step (the CPU path runs per "
                    "image and sums, BASELINE.md 4.4); rank 0 only"}
    legs = set(args.legs.split(",")) if args.legs else set()
    sub_steps, sub_warm = max(1, min(steps, 5)), min(warmup, 3)
    if "train" in legs:
        tv, _, tsample, tper = cpu_train_leg(sub_steps, sub_warm)
        line["train"] = {"value": tv, "unit": UNIT, "config": {"workload": TR_WORKLOAD}, "sample": tsample, "s_per_step_sample": tper}
    if "meet_gqa" in legs:
        mv, _, msample = cpu_meet_leg(sub_steps, sub_warm)
        line["meet_gqa"] = {"value": mv, "unit": UNIT, "config": {"workload": MEET_WORKLOAD}, "sample": msample}
    if "sweep96" in legs:
        pv, _, psample, pper, _ = cpu_infer_leg(380, sub_steps, sub_warm, n_boxes=TR_BOXES, mode="predcls")
        n_p, n_s = 48 * 380, 48 * N_BOXES * (N_BOXES - 1)
        total_s = n_p * pper + n_s * per_pair
        line["sweep96"] = {"value": (n_p + n_s) / total_s, "unit": UNIT, "config": {"workload": SWEEP_WORKLOAD},
                           "sample": "per-image times summed: 48 x (" + psample + ") + 48 x (headline sample)"}
    line["wall_s"] = time.perf_counter() - t0
    emit(line)


def _json_only_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner on stdout
    when NCCL_DEBUG is set in the environment), so keep a private handle on the real stdout for the line and point
    file descriptor 1 at stderr for everything else."""
    global _OUT
    if _OUT is None:
        sys.stdout.flush()
        _OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)
    return _OUT


_OUT = None


def emit(line):
    out = _json_only_stdout()
    out.write(json.dumps(line) + "\n")
    out.flush()


# ------------------------------------------------------------------------------------------------ GPU arm
def main():
    args = parse()
    _json_only_stdout()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from veto_b200 import lib as L
    from veto_b200 import ops, registry, synth
    from veto_b200 import workloads as WL
    from veto_b200.distributed import allreduce_flat, allreduce_gradients, finish_gradient_sync, max_over_ranks, shard_images
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
    legs = set(args.legs.split(",")) if args.legs else set()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)" if peaks else "1.4 PFLOP/s sustained (of fallback)"
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    # tensor-pipe cost of one product in bf16-MMA equivalents (kind::f16 = 1, kind::f8f6f4 = 0.5 per K element)
    units = {"bf16x3": 3, "f16c8": 2, "bf16": 1, "f16": 1, "fp32": 1}[args.precision]
    train_precision = L.TRAIN_PRECISION[args.precision]
    train_units = {"bf16x3": 3, "bf16": 1, "fp32": 1}[train_precision]
    pin = lambda t: t.detach().cpu().pin_memory()
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    cpu_ok = rank == 0 and world == 1 and not args.no_cpu_baseline

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

    def nbytes(ts):
        return sum(t.numel() * t.element_size() for t in ts)

    # ------------------------------------------------------------------------------------------ inference legs
    def build_head(mode, precision, predictor="VETOPredictor", dataset="VG", state=None):
        cfg = WL.make_cfg(predictor=predictor, mode=mode, dataset=dataset, max_pairs=MAX_PAIRS, precision=precision,
                          chunk_pairs=args.chunk)
        pred = WL.build_predictor(cfg, state if state is not None else synth.predictor_state(11), dev)
        fe = registry.make_roi_box_feature_extractor(cfg, 256, for_relation=True).to(dev).eval()
        samp = make_roi_relation_samp_processor(cfg)
        post = make_roi_relation_post_processor(cfg)

        def head(feats, depth, bls):
            """ROIRelationHead.forward at test time (relation_head.py:134-243)."""
            pairs = samp.prepare_test_pairs(dev, bls)
            x2d, d2d, _, _ = fe(feats, bls, depth_features=depth)
            out = pred(bls, pairs, None, None, roi_features=x2d, roi_depth_features=d2d)
            return post((out[1], [b.get_field("predict_logits") for b in bls]), pairs, bls, incre_idx_list=out[3])

        return head, pred

    def inference_leg(name, workload, n_boxes_list, mode, predictor="VETOPredictor", dataset="VG", state=None, num_obj=151,
                      steps=None, cpu=None):
        """One inference workload: resident timing, e2e timing from pinned host buffers, per-stage device times."""
        Bi = len(n_boxes_list)
        Ri = sum(n * (n - 1) for n in n_boxes_list)
        batch = synth.make_batch(200 + rank, n_boxes_list, H=IMG_H, W=IMG_W, num_obj=num_obj, mode=mode, features=False)
        if mode == "sgdet":
            synth.add_nms_fields(batch, 300 + rank, relabel=False)       # the detector's per-class boxes (late NMS input)
        feats, depth = WL.random_features(Bi, IMG_H, IMG_W, dev, g)
        head, pred = build_head(mode, args.precision, predictor, dataset, state)
        bls = WL.boxlists(batch, dev, num_obj)

        def resident():
            with torch.no_grad():
                return head(feats, depth, bls)

        feats_host, depth_host = [pin(f) for f in feats], pin(depth)
        boxes_host, fields_host = host_boxlists(bls)
        h2d = nbytes(feats_host + [depth_host] + boxes_host) + sum(nbytes(f.values()) for f in fields_host)
        res0 = resident()
        out_rows = sum(int(r.get_field("pred_rel_scores").shape[0]) for r in res0)
        n_cols = int(res0[0].get_field("pred_rel_scores").shape[1])
        probs_host = torch.empty((out_rows, n_cols), dtype=torch.float32).pin_memory()      # ranked class probabilities
        pair_dtype = res0[0].get_field("rel_pair_idxs").dtype
        pairs_host = torch.empty((out_rows, 2), dtype=pair_dtype).pin_memory()              # ranked pairs
        del res0

        def slot():
            f = [dev_like(t) for t in feats_host]
            d = dev_like(depth_host)
            b, copies = boxlist_slot(boxes_host, fields_host)
            copies += list(zip(f, feats_host)) + [(d, depth_host)]
            return (f, d, b), copies

        pipe = InputPipe(slot)

        def e2e():
            with torch.no_grad():
                sl = pipe.acquire()
                f, d, b = sl["struct"]
                res = head(f, d, b)
                pipe.release(sl)
                probs_host.copy_(torch.cat([r.get_field("pred_rel_scores") for r in res]), non_blocking=True)
                pairs_host.copy_(torch.cat([r.get_field("rel_pair_idxs") for r in res]), non_blocking=True)

        k = steps or args.steps
        with ClockSampler(local) as clk:
            ms, launches = timed(resident, k, args.warmup)
        clocks = clk.summary()
        ms_e2e, _ = timed(e2e, max(2, min(k, 5)), 2)
        torch.cuda.synchronize()
        with ops.StageTimer() as st:
            resident()
        M = Ri * 19
        gemm_tags = ["gemm_qkv", "gemm_out", "gemm_ff1", "gemm_ff2"]
        g_ms = sum(st.ms.get(t, 0.0) for t in gemm_tags)
        g_launch = sum(st.launches.get(t, 0) for t in gemm_tags)
        # FLOPs of the encoder Linears as EXECUTED: five full layers; the last layer computes K, V for every token and
        # everything else for the CLS row only (model_veto.py:25 consumes x[:,0] alone)
        flops_exec = 5 * 2.0 * M * ENC_MAC_PER_ROW + 2.0 * M * 1152 * 576 + 2.0 * Ri * (576 * 576 * 2 + 2 * 576 * 1152)
        flops_ref = 6 * 2.0 * M * ENC_MAC_PER_ROW          # the reference computes all six layers in full
        ach = flops_exec / (g_ms * 1e-3) / 1e12 if g_ms > 0 else 0.0
        total_ms = sum(st.ms.values())
        nb = sum(n_boxes_list)
        hbm_bytes = {
            "pairs": Ri * (16 + 16 + 8),
            "roi_gather": nb * 131072,                                       # outputs only; map reads are L2-resident per image
            "tokens": Ri * 19 * 576 * 4,
            "layernorm": (2 * 5 * M + M + Ri) * 4608,
            "attention": 5 * M * (6912 + 2304) + M * 4608 + Ri * 4608,
            "postprocess": Ri * (n_cols * 4 * 2 + 16 * 2 + 8 + 4),
        }
        hbm = {t: {"ms": round(st.ms[t], 3), "launches": st.launches.get(t), "algorithmic_bytes": int(v),
                   "achieved_gbs": round(v / st.ms[t] / 1e6, 1), "frac_of_measured_hbm": round(v / st.ms[t] / 1e6 / hbm_peak, 3)}
               for t, v in hbm_bytes.items() if st.ms.get(t)}
        traffic, traffic_src = None, None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "gemm_traffic.json")))
            traffic, traffic_src = tj.get("infer_" + args.precision), tj.get("infer_source")
        except Exception:
            pass
        roofline = {
            "bound": "tensor",
            "kernel": ("gemm_tc2_kernel (tcgen05.mma cta_group::2 kind::f16 / kind::f8f6f4, TMEM accumulators, TMA): the encoder Linears"
                       if args.precision != "fp32" else "gemm_simt_kernel"),
            "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf, "traffic": traffic,
            "peak_source": peak_src, "traffic_source": traffic_src,
            "algorithmic_flops_per_launch": flops_exec / max(g_launch, 1), "launches_per_step": g_launch,
            "avg_launch_ms": g_ms / max(g_launch, 1), "share_of_step": g_ms / total_ms if total_ms else None,
            "mma_units_per_product": units, "executed_mma_tflops": ach * units, "frac_executed": ach * units / peak_tf,
            "reference_formulation_tflops": flops_ref / (g_ms * 1e-3) / 1e12 if g_ms > 0 else 0.0,
            "note": ("achieved = 2*M*N*K of the encoder Linears the launches execute (the last layer is evaluated for the CLS row "
                     "only) / CUDA-event time of those launches; one product costs `mma_units_per_product` bf16-MMA equivalents "
                     "(f16c8: fp16 product + e4m3 correction over 2K)"),
            "stage_ms": {t: round(v, 3) for t, v in st.ms.items()}, "stage_launches": dict(st.launches),
        }
        cpu_obj = None
        if cpu_ok and cpu is not None:
            cpu_obj = cpu()
        leg = {
            "metric": METRIC, "value": world * Ri / ms * 1e3, "unit": UNIT, "ms_per_step": ms, "steps": k, "warmup": args.warmup,
            "images_per_sec": world * Bi / ms * 1e3,
            "config": {"workload": workload, "images_per_gpu": Bi, "pairs_per_step_per_gpu": Ri, "precision": args.precision,
                       "chunk_pairs": pred.chunk_pairs if hasattr(pred, "chunk_pairs") and pred.chunk_pairs else "library default (7976)"},
            "e2e": {"value": world * Ri / ms_e2e * 1e3, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": nbytes([probs_host, pairs_host]), "ms_per_step": ms_e2e,
                    "pipeline": "double-buffered H2D on a copy stream, overlapped with the previous step"},
            "gpu_launches": launches, "tflops_reference_formulation": world * Ri / ms * 1e3 * FLOP_PER_PAIR / 1e12,
            "roofline": roofline, "hbm_kernels": hbm, "cpu_baseline": cpu_obj, "clocks": clocks,
        }
        del pipe, feats, depth, feats_host, depth_host, pred, head
        ops._workspaces.clear()
        torch.cuda.empty_cache()
        return leg

    def cpu_headline():
        v, cores, sample, _, _ = cpu_infer_leg(args.cpu_sample_pairs, 5, 3)
        return {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}

    headline = inference_leg("inference", INF_WORKLOAD, [N_BOXES] * N_IMAGES, "sgdet", cpu=cpu_headline)

    # ------------------------------------------------------------------------------------------ configs[3]
    meet_leg = None
    if "meet_gqa" in legs:
        def cpu_meet():
            v, cores, sample = cpu_meet_leg(5, 3)
            return {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
        mstate = synth.meet_state(14, 201, synth.GROUP_SPLITS[("GQA", "divide4")])
        meet_leg = inference_leg("meet_gqa", MEET_WORKLOAD, [MEET_BOXES] * MEET_IMAGES, "predcls", predictor="VETOPredictor_MEET",
                                 dataset="GQA", state=mstate, num_obj=201, cpu=cpu_meet)
        meet_leg.pop("clocks", None)

    # ------------------------------------------------------------------------------------------ configs[4]
    sweep_leg = None
    if "sweep96" in legs:
        n_all = [TR_BOXES] * 48 + [N_BOXES] * 48                     # images 0..47 PredCls, 48..95 SGDet
        mine = shard_images(n_all, world, MAX_PAIRS)[rank]           # pair-count balanced (LPT), the same table on every rank
        mine_p, mine_s = [i for i in mine if i < 48], [i for i in mine if i >= 48]
        sw = {}
        for tag, idx, mode, nbx in (("p", mine_p, "predcls", TR_BOXES), ("s", mine_s, "sgdet", N_BOXES)):
            if not idx:
                continue
            b = synth.make_batch(400 + rank, [nbx] * len(idx), H=IMG_H, W=IMG_W, mode=mode, features=False)
            if mode == "sgdet":
                synth.add_nms_fields(b, 500 + rank, relabel=False)
            f, d = WL.random_features(len(idx), IMG_H, IMG_W, dev, g)
            head, _ = build_head(mode, args.precision)
            sw[tag] = (head, f, d, WL.boxlists(b, dev, 151))

        def sweep_step():
            with torch.no_grad():
                for head, f, d, bl in sw.values():
                    head(f, d, bl)

        total_pairs = sum(n * (n - 1) for n in n_all)
        sweep_ms, sweep_launch = timed(sweep_step, max(2, min(args.steps, 5)), 2)
        my_pairs = sum(n_all[i] * (n_all[i] - 1) for i in mine)
        scpu = None
        if cpu_ok:
            pv, cores, psample, pper, _ = cpu_infer_leg(380, 5, 3, n_boxes=TR_BOXES, mode="predcls")
            per_s = 1.0 / headline["cpu_baseline"]["value"] if headline.get("cpu_baseline") else None
            if per_s:
                n_p, n_s = 48 * 380, 48 * N_BOXES * (N_BOXES - 1)
                scpu = {"value": (n_p + n_s) / (n_p * pper + n_s * per_s), "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": "per-image times summed: 48 x (" + psample + ") + 48 x (the headline's sample)"}
        sweep_leg = {"metric": METRIC, "value": total_pairs / sweep_ms * 1e3, "unit": UNIT, "ms_per_step": sweep_ms,
                     "images_per_sec": 96 / sweep_ms * 1e3, "scaling": "strong",
                     "config": {"workload": SWEEP_WORKLOAD, "images_total": 96, "pairs_total": total_pairs,
                                "images_this_rank": len(mine), "pairs_this_rank": my_pairs, "precision": args.precision,
                                "sharding": f"veto_b200.distributed.shard_images over {world} rank(s), no collective"},
                     "gpu_launches": sweep_launch, "cpu_baseline": scpu}
        del sw
        ops._workspaces.clear()
        torch.cuda.empty_cache()

    # ------------------------------------------------------------------------------------------ configs[1]: training
    train_leg = None
    if "train" in legs:
        from veto_b200 import depth_backbone as DB
        B = TR_IMAGES
        R = B * TR_BOXES * (TR_BOXES - 1)
        batch = synth.make_batch(100 + rank, [TR_BOXES] * B, H=IMG_H, W=IMG_W, mode="predcls", features=False)
        feats_dev, _ = WL.random_features(B, IMG_H, IMG_W, dev, g)
        dimg_dev = torch.from_numpy(synth.depth_images(B, IMG_H, IMG_W, seed=rank)).to(dev)
        state = synth.predictor_state(11, spread=False)                   # torch-default-like init: a sane training start
        cfg = WL.make_cfg(mode="predcls", precision=args.precision)
        pred = registry.make_roi_relation_predictor(cfg, 512)
        pred.load_state_dict(synth.to_torch_state(state), strict=True)
        pred = pred.to(dev).train()
        dmodel = DB.build_resnet18_depth(cfg).to(dev).train()
        dmodel.load_state_dict({k: torch.from_numpy(v) for k, v in synth.depth_state(0).items()})
        fe = registry.make_roi_box_feature_extractor(cfg, 256, for_relation=True).to(dev).train()
        samp = make_roi_relation_samp_processor(cfg)
        bls_dev = WL.boxlists(batch, dev, 151)
        rel_mats_np = synth.make_relation_matrices(7 + rank, [TR_BOXES] * B, 51, 10)   # ~10 annotated relations per image

        def make_targets(bls, mats):
            out = []
            for bl, m in zip(bls, mats):
                t = BoxList(bl.bbox, (IMG_W, IMG_H), "xyxy")
                t.add_field("relation", m)
                out.append(t)
            return out

        targets_dev = make_targets(bls_dev, [torch.from_numpy(m).to(dev) for m in rel_mats_np])
        head_params = [p for p in pred.parameters() if p.requires_grad]
        depth_params = [p for p in dmodel.parameters() if p.requires_grad]
        params = head_params + depth_params                               # the reference's train_modules (relation_train_net.py:166-170)
        opt = torch.optim.Adam(params, lr=1e-4 * B, fused=True)          # BASE_LR x IMS_PER_BATCH (:330-339)
        losses = []
        sync = {"works": []}
        # the relation head's gradients are complete when its loss leaves the predictor (veto_relation_train_step computes
        # them with the loss): their all-reduce starts there and runs under the ROIAlign + depth-backbone backward
        pred.grad_sync = (lambda flat: sync["works"].extend(allreduce_flat(flat))) if world > 1 else None

        def train_step(feats, dimg, bls, targets, pending=None):
            opt.zero_grad(set_to_none=True)
            pend = pending if pending is not None else samp.gtbox_relsample_async(bls, targets)
            _, rel_labels, pairs, _ = pend.result()
            depth = dmodel(dimg)                                          # generalized_rcnn.py:53-54
            x2d, d2d, _, _ = fe(feats, bls, depth_features=depth)
            loss = pred(bls, pairs, rel_labels, None, roi_features=x2d, roi_depth_features=d2d)[2]["rel_loss"]
            loss.backward()
            allreduce_gradients(depth_params)                             # the backbone's flat gradient buffer, in place
            finish_gradient_sync(sync["works"])
            sync["works"].clear()
            torch.nn.utils.clip_grad_norm_(params, 5.0, foreach=True)     # GRAD_NORM_CLIP 5.0 (:475-481)
            opt.step()
            return loss.detach()

        resident = {"pending": None}

        def step_resident():
            pend = resident["pending"] or samp.gtbox_relsample_async(bls_dev, targets_dev)
            resident["pending"] = samp.gtbox_relsample_async(bls_dev, targets_dev)   # the NEXT step's sampling
            losses.append(train_step(feats_dev, dimg_dev, bls_dev, targets_dev, pend))

        feats_host = [pin(f) for f in feats_dev]
        dimg_host = pin(dimg_dev)
        boxes_host, fields_host = host_boxlists(bls_dev)
        labels_host = [pin(t.get_field("relation")) for t in targets_dev]
        h2d_bytes = nbytes(feats_host + [dimg_host] + boxes_host + labels_host) + sum(nbytes(f.values()) for f in fields_host)
        loss_host = torch.empty(1, dtype=torch.float32).pin_memory()

        def train_slot():
            feats = [dev_like(f) for f in feats_host]
            dimg = dev_like(dimg_host)
            bls, copies = boxlist_slot(boxes_host, fields_host)
            mats = [dev_like(l) for l in labels_host]
            targets = make_targets(bls, mats)
            copies += list(zip(feats, feats_host)) + [(dimg, dimg_host)] + list(zip(mats, labels_host))
            return (feats, dimg, bls, targets), copies

        def sample_with_inputs(sl):
            pend = samp.gtbox_relsample_async(sl["struct"][2], sl["struct"][3])
            for t in [pend.pairs, pend.labels] + list(pend.binaries):
                t.record_stream(torch.cuda.default_stream(dev))
            return pend

        train_pipe = InputPipe(train_slot, after_small=sample_with_inputs)

        def step_e2e():
            sl = train_pipe.acquire()
            feats, dimg, bls, targets = sl["struct"]
            pend, sl["pending"] = sl["pending"], None
            loss = train_step(feats, dimg, bls, targets, pend)
            train_pipe.release(sl)
            loss_host.copy_(loss.reshape(1), non_blocking=True)

        ms_step, launches = timed(step_resident, args.steps, args.warmup)
        loss_first, loss_last = float(losses[0]), float(losses[-1])
        ms_e2e, _ = timed(step_e2e, max(2, min(args.steps, 5)), 2)
        torch.cuda.synchronize()
        with ops.StageTimer() as st:
            step_resident()
        M = R * 19
        enc_flops = 6 * 2.0 * M * ENC_MAC_PER_ROW
        fwd_tags = ["gemm_qkv", "gemm_out", "gemm_ff1", "gemm_ff2"]
        g_ms = sum(st.ms.get(k, 0.0) for k in fwd_tags + ["bwd_dgrad", "bwd_wgrad"])
        g_launch = sum(st.launches.get(k, 0) for k in fwd_tags + ["bwd_dgrad", "bwd_wgrad"])
        algo_flops = 3.0 * enc_flops                                           # forward + input-gradient + weight-gradient
        achieved = algo_flops / (g_ms * 1e-3) / 1e12 if g_ms > 0 else 0.0
        total_stage_ms = sum(st.ms.values())
        traffic, traffic_src = None, None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "gemm_traffic.json")))
            traffic, traffic_src = tj.get("train_" + train_precision), tj.get("train_source")
        except Exception:
            pass
        # the gradient exchange alone (N > 1): both flat buffers, timed back to back
        comm = None
        if world > 1:
            def comm_only():
                finish_gradient_sync(allreduce_gradients(params, async_op=True))
            comm_ms, _ = timed(comm_only, 10, 3)
            comm = {"isolated_ms": comm_ms, "bytes": int(sum(p.grad.numel() for p in params if p.grad is not None) * 4),
                    "note": "in-place NCCL all-reduce (AVG) of the two flat gradient buffers; in the step the head's 70 MB start "
                            "when its loss is computed and overlap the ROIAlign + depth-backbone backward"}
        tcpu = None
        if cpu_ok:
            v, cores, sample, _ = cpu_train_leg(5, 3)
            tcpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
        train_ws_gb = sum(t.numel() for t in ops._workspaces.values()) / 1e9
        stage_ms = {k: round(v, 3) for k, v in st.ms.items()}
        if comm:
            stage_ms["comm_allreduce_isolated"] = round(comm["isolated_ms"], 3)
        train_leg = {
            "metric": METRIC, "value": world * R / ms_step * 1e3, "unit": UNIT, "ms_per_step": ms_step, "steps": args.steps,
            "warmup": args.warmup, "images_per_sec": world * B / ms_step * 1e3, "scaling": "weak",
            "dtype": train_precision,
            "config": {"workload": TR_WORKLOAD, "images_per_gpu": B, "pairs_per_step_per_gpu": R, "precision": train_precision,
                       "parallelism": f"image-sharded data parallel x{world}; in-place NCCL all-reduce of the flat gradient buffers "
                                      "(head: overlapped with the backbone backward)",
                       "optimizer": "Adam (torch fused) over the relation head + depth backbone, clip_grad_norm 5.0"},
            "tflops_reference_formulation": world * R / ms_step * 1e3 * 3 * FLOP_PER_PAIR / 1e12,
            "loss_first_last": [loss_first, loss_last], "train_workspace_gb": round(train_ws_gb, 2),
            "e2e": {"value": world * R / ms_e2e * 1e3, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e,
                    "pipeline": "double-buffered: the H2D copy of step k+1 (copy stream) overlaps the training of step k"},
            "gpu_launches": launches,
            "roofline": {"bound": "tensor", "kernel": "gemm_tc2_kernel / gemm_tn2_kernel: forward + dX, and dW GEMMs of the encoder",
                         "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf, "traffic": traffic,
                         "peak_source": peak_src, "traffic_source": traffic_src,
                         "algorithmic_flops_per_launch": algo_flops / max(g_launch, 1), "launches_per_step": g_launch,
                         "avg_launch_ms": g_ms / max(g_launch, 1),
                         "share_of_step": g_ms / total_stage_ms if total_stage_ms else None,
                         "mma_units_per_product": train_units, "executed_mma_tflops": achieved * train_units,
                         "frac_executed": achieved * train_units / peak_tf, "stage_ms": stage_ms,
                         "stage_launches": dict(st.launches)},
            "comm": comm, "cpu_baseline": tcpu,
        }
        resident["pending"] = None
        del opt, params, pred, dmodel, fe, feats_dev, feats_host, train_pipe
        ops._workspaces.clear()
        torch.cuda.empty_cache()

    if rank == 0:
        line = {
            "metric": METRIC, "value": headline["value"], "unit": UNIT, "n_gpus": world, "steps": headline["steps"],
            "warmup": args.warmup, "ms_per_step": headline["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None,
            "dtype": {"bf16x3": "bf16x3 (split-bf16 tensor-core products, fp32 accumulate; fp32-grade)", "bf16": "bf16",
                      "f16c8": "f16c8 (fp16 tensor-core product + e4m3 first-order corrections, fp32 accumulate; within 1e-3 of fp32)",
                      "f16": "f16", "fp32": "f32"}[args.precision],
            "data": "synthetic", "config": dict(HEADLINE_CONFIG), "precision": args.precision,
            "parallelism": f"image-sharded data parallel x{world}, no collective on the inference path",
            "images_per_sec": headline["images_per_sec"],
            "tflops_reference_formulation": headline["tflops_reference_formulation"],
            "e2e": headline["e2e"], "gpu_launches": headline["gpu_launches"], "clocks": headline["clocks"],
            "roofline": headline["roofline"], "hbm_peak_gbs": hbm_peak, "hbm_kernels": headline["hbm_kernels"],
            "cpu_baseline": headline["cpu_baseline"],
            "train": train_leg, "meet_gqa": meet_leg, "sweep96": sweep_leg,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
