"""Micro-benchmark of the weight-gradient GEMM (gemm_tn2) at the training shapes (run on the GPU box; not a pytest file)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from veto_b200 import ops
DUMP = "/tmp/veto_tn_dump.txt"
os.environ["VETO_PROFILE_DUMP"] = DUMP
dev = torch.device("cuda:0")
rows = 86640
torch.manual_seed(0)
for Nw, Kw in ((1728, 576), (576, 576), (1152, 576), (576, 1152)):
    y = torch.randn(rows, Nw, device=dev)
    x = torch.randn(rows, Kw, device=dev)
    line = []
    for split in (3, 4, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16):
        ops.test_gemm_tn(y, x, "bf16x3", split, None)
        torch.cuda.synchronize()
        with ops.StageTimer() as st:
            for _ in range(3):
                ops.test_gemm_tn(y, x, "bf16x3", split, None)
        # per-launch times of the profiled calls (VETO_PROFILE_DUMP): convert(y), convert(x), GEMM[, split-K reduce] per call
        us = [float(l.split()[2]) for l in open(DUMP).read().splitlines()]
        os.remove(DUMP)
        per = len(us) // 3
        gemm = sorted(us[i * per + 2] for i in range(3))[1]
        line.append(f"{split}:{gemm:.0f}")
    flops = 2.0 * rows * Nw * Kw * 3
    print(f"Nw={Nw} Kw={Kw} us by split-K: " + " ".join(line) + f"   (1.4 PF/s floor {flops / 1.4e15 * 1e6:.0f} us)", flush=True)
