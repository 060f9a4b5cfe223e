#!/usr/bin/env python
"""bench.py — VETO relation-head throughput on B200 (the metric of BASELINE.json), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--precision bf16x3|bf16|fp32]

Workload: BASELINE.json configs[2] — SGDet-shaped inference, 80 proposals / image (6320 candidate pairs, cap
MAX_PROPOSAL_PAIR 8192), batch 32, VG 151/51, one batch per GPU.  (configs[1] is a training step; the training
branch of the head is not built yet, so the largest single-GPU inference configuration is the bench line — DESIGN.md.)
A "step" = prepare_test_pairs -> VETOFeatureExtractor (ROI gather) -> VETOPredictor forward for one batch.

  value : relation pairs / s, whole job, inputs resident in HBM, CUDA-event timed, max over ranks.
  e2e   : the same metric through the public API from pinned HOST buffers: H2D of the step's feature maps, boxes
          and logits and D2H of the relation logits inside the timed region.
  roofline : tcgen05 GEMM kernel (the dominant one): algorithmic FLOPs per launch / CUDA-event duration per
          launch (veto_profile_*), against MEASURED_PEAKS.json bf16_tflops_sustained.
  cpu_baseline : the numpy/C oracle (port of the reference's CPU path) on a bounded sample, host cores.

Multi-GPU: images are independent, so every rank runs its own batch (weak scaling, no data-path collective).
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
N_IMAGES, N_BOXES, MAX_PAIRS = 32, 80, 8192
IMG_H, IMG_W = 592, 800
FLOP_PER_PAIR = 648.7e6          # reference formulation (SURVEY.md §8d / BASELINE.md §3)
WORKLOAD = ("configs[2]: SGDet-shaped inference, 32 images x 80 proposals (6320 pairs/image, MAX_PROPOSAL_PAIR 8192), "
            "VG 151/51, 592x800 images, P2-P5 + depth NCHW fp32")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default=os.environ.get("VETO_PRECISION", "bf16x3"), choices=["bf16x3", "bf16", "fp32"])
    ap.add_argument("--chunk", type=int, default=int(os.environ.get("VETO_CHUNK_PAIRS", "0")))
    ap.add_argument("--images", type=int, default=N_IMAGES)
    ap.add_argument("--cpu-sample-pairs", type=int, default=int(os.environ.get("VETO_CPU_SAMPLE_PAIRS", "2048")))
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
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


# ------------------------------------------------------------------------------------------------ CPU baseline
def cpu_reference_leg(sample_pairs: int, steps: int = 1, warmup: int = 0):
    """The reference's CPU path for this workload, restated by the oracle (numpy + C ROIAlign; the Python reference
    itself cannot travel to the GPU box): one image of the bench shape — pair enumeration for 80 boxes, ROI gather of
    the 80 boxes, and the predictor on the first `sample_pairs` of its 6320 pairs.  Returns (pairs/s, cores, sample)."""
    import torch
    from oracle import torch_port as TP
    from veto_b200 import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    batch = synth.make_batch(1000, [N_BOXES], H=IMG_H, W=IMG_W, mode="sgdet")
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
            t_pairs = time.perf_counter() - t1
            assert bool(torch.isfinite(logits).all())
            if it >= warmup:
                # per-image fixed cost amortised over the image's 6320 pairs + per-pair cost of the sampled pairs
                times.append(t_fixed / len(pairs[0]) + t_pairs / len(sub[0]))
    per_pair = sorted(times)[len(times) // 2]
    sample = (f"1 image x {N_BOXES} proposals: pair enumeration + ROI gather of {N_BOXES} boxes (amortised over 6320 pairs) "
              f"+ predictor on the first {len(sub[0])} pairs (the reference materialises 262 KB per pair, 6320 at once "
              f"need 1.66 GB); oracle/torch_port.py = the reference's formulation on torch CPU fp32 kernels, {cores} threads")
    return 1.0 / per_pair, cores, sample, per_pair


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.perf_counter()
    val, cores, sample, per_pair = cpu_reference_leg(args.cpu_sample_pairs, steps=max(1, min(args.steps, 3)),
                                                     warmup=min(args.warmup, 1))
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": per_pair * args.cpu_sample_pairs * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "l2": "host CPU run"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": time.perf_counter() - t0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    from tests import harness as H
    from veto_b200 import lib as L
    from veto_b200 import ops, registry, synth
    from veto_b200.distributed import max_over_ranks
    from veto_b200.sampling import make_roi_relation_samp_processor

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L.require_device()

    # ---- synthetic batch of this rank (independent images: every rank has its own, weak scaling)
    B = args.images
    batch = synth.make_batch(100 + rank, [N_BOXES] * B, H=IMG_H, W=IMG_W, mode="sgdet", features=False)
    rgb_hw, depth_hw = synth.fpn_shapes(IMG_H, IMG_W)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    feats_dev = [torch.randn((B, 256) + hw, generator=g, device=dev) for hw in rgb_hw]
    depth_dev = torch.relu(torch.randn((B, 256) + depth_hw, generator=g, device=dev))
    state = synth.predictor_state(11)
    cfg = H.make_cfg(mode="sgdet", max_pairs=MAX_PAIRS, precision=args.precision, chunk_pairs=args.chunk)
    pred = H.build_predictor(cfg, state, dev)
    fe = registry.make_roi_box_feature_extractor(cfg, 256, for_relation=True).to(dev).eval()
    samp = make_roi_relation_samp_processor(cfg)
    bls_dev = H.boxlists(batch, dev, 151)
    R = B * N_BOXES * (N_BOXES - 1)

    def step_resident():
        with torch.no_grad():
            pairs = samp.prepare_test_pairs(dev, bls_dev)
            x2d, d2d, _, _ = fe(feats_dev, bls_dev, depth_features=depth_dev)
            out = pred(bls_dev, pairs, None, None, roi_features=x2d, roi_depth_features=d2d)
        return out[1]

    # ---- host-side copies for the end-to-end leg (pinned)
    pin = lambda t: t.cpu().pin_memory()
    feats_host = [pin(f) for f in feats_dev]
    depth_host = pin(depth_dev)
    boxes_host = [pin(b.bbox) for b in bls_dev]
    fields_host = [{k: pin(b.get_field(k)) for k in ("labels", "predict_logits", "pred_scores", "pred_labels")} for b in bls_dev]
    h2d_bytes = sum(t.numel() * t.element_size() for t in feats_host + [depth_host] + boxes_host)
    h2d_bytes += sum(t.numel() * t.element_size() for f in fields_host for t in f.values())
    logits_host = torch.empty((R, 51), dtype=torch.float32).pin_memory()
    d2h_bytes = logits_host.numel() * 4
    from veto_b200.structures import BoxList

    def step_e2e():
        with torch.no_grad():
            feats = [f.to(dev, non_blocking=True) for f in feats_host]
            depth = depth_host.to(dev, non_blocking=True)
            bls = []
            for bb, ff in zip(boxes_host, fields_host):
                bl = BoxList(bb.to(dev, non_blocking=True), (IMG_W, IMG_H), "xyxy")
                for k, v in ff.items():
                    bl.add_field(k, v.to(dev, non_blocking=True))
                bls.append(bl)
            pairs = samp.prepare_test_pairs(dev, bls)
            x2d, d2d, _, _ = fe(feats, bls, depth_features=depth)
            rel = pred(bls, pairs, None, None, roi_features=x2d, roi_depth_features=d2d)[1]
            logits_host.copy_(torch.cat(list(rel)), non_blocking=True)

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

    # ---- headline: resident inputs
    with ClockSampler(local) as clk:
        ms_step, launches = timed(step_resident, args.steps, args.warmup)
    clocks = clk.summary()
    value = world * R / ms_step * 1e3
    # ---- end to end from host buffers
    ms_e2e, _ = timed(step_e2e, max(1, min(args.steps, 3)), 1)
    e2e_value = world * R / ms_e2e * 1e3

    # ---- per-stage device time of one step (CUDA events around every launch) -> roofline of the GEMM kernel
    torch.cuda.synchronize()
    with ops.StageTimer() as st:
        step_resident()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)" if peaks else "1.4 PFLOP/s sustained (of fallback)"
    M = R * 19
    gemm_flops = {"gemm_qkv": 2.0 * M * 1728 * 576, "gemm_out": 2.0 * M * 576 * 576, "gemm_ff1": 2.0 * M * 1152 * 576,
                  "gemm_ff2": 2.0 * M * 576 * 1152}
    layers = 6
    g_ms = sum(st.ms.get(k, 0.0) for k in gemm_flops)
    g_launch = sum(st.launches.get(k, 0) for k in gemm_flops)
    algo_flops = layers * sum(gemm_flops.values())
    passes = {"bf16x3": 3, "bf16": 1, "fp32": 1}[args.precision]
    achieved = algo_flops / (g_ms * 1e-3) / 1e12 if g_ms > 0 else 0.0
    total_stage_ms = sum(st.ms.values())
    traffic, traffic_src = None, None
    try:  # DRAM bytes per launch of the GEMM kernel from the committed ncu --set full capture (profiles/)
        tj = json.load(open(os.path.join(ROOT, "profiles", "gemm_traffic.json")))
        traffic, traffic_src = tj.get(args.precision), tj.get("source")
    except Exception:
        pass
    roofline = {
        "bound": "tensor", "kernel": "gemm_tc2_kernel (tcgen05.mma cta_group::2, TMEM accumulators, TMA)" if args.precision != "fp32" else "gemm_simt_kernel",
        "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf, "traffic": traffic,
        "peak_source": peak_src, "traffic_source": traffic_src,
        "algorithmic_flops_per_launch": algo_flops / max(g_launch, 1), "launches_per_step": g_launch,
        "avg_launch_ms": g_ms / max(g_launch, 1), "share_of_step": g_ms / total_stage_ms if total_stage_ms else None,
        "executed_mma_tflops": achieved * passes, "frac_executed": achieved * passes / peak_tf,
        "note": "algorithmic = 2*M*N*K of the reference's fp32 Linear layers; bf16x3 executes 3 bf16 MMAs per product",
        "stage_ms": {k: round(v, 3) for k, v in st.ms.items()},
    }

    # ---- HBM-bound kernels of the path: algorithmic bytes per step / event time, against the measured copy bandwidth
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    n_boxes_total = B * N_BOXES
    rows_full = (layers - 1) * M                      # token rows the 5 full layers normalise / attend
    hbm_bytes = {
        "pairs": R * (16 + 16 + 8),                   # enumerate writes 16 B/pair; globalize reads 16, writes 8
        "roi_gather": n_boxes_total * 131072,         # bytes written (each touched map element is read once on top)
        "tokens": R * 19 * 576 * 4,                   # token block written; box-level tables are L2-resident
        "layernorm": (2 * rows_full + M + R) * 4608,  # 2304 B read + 2304 B written per row (last layer: LN2 on CLS rows)
        "attention": rows_full * (6912 + 2304) + M * 4608 + R * 4608,  # qkv read + hi/lo output written
    }
    hbm_kernels = {k: {"ms": round(st.ms[k], 3), "algorithmic_bytes": int(v), "achieved_gbs": round(v / st.ms[k] / 1e6, 1),
                       "frac_of_measured_hbm": round(v / st.ms[k] / 1e6 / hbm_peak, 3)}
                   for k, v in hbm_bytes.items() if st.ms.get(k)}

    # ---- CPU baseline on rank 0 (bounded sample)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, cores, sample, _ = cpu_reference_leg(args.cpu_sample_pairs)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"bf16x3": "bf16x3 (split-bf16 tensor-core products, fp32 accumulate; fp32-grade)", "bf16": "bf16",
                      "fp32": "f32"}[args.precision],
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "images_per_gpu": B, "pairs_per_step_per_gpu": R, "precision": args.precision,
                       "chunk_pairs": pred.chunk_pairs or "library default", "parallelism": f"image-sharded x{world}",
                       "l2": "inputs larger than L2 (1.38 GB of feature maps + 0.34 GB of ROI features per step)"},
            "images_per_sec": world * B / ms_step * 1e3,
            "tflops_reference_formulation": value * FLOP_PER_PAIR / 1e12,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                    "ms_per_step": ms_e2e},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "hbm_kernels": hbm_kernels,
            "hbm_peak_gbs": hbm_peak, "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
