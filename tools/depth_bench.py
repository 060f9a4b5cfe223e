"""Times the depth backbone (f3) forward / forward+backward on a BASELINE-size batch with CUDA events."""
import json
import sys

import torch

sys.path.insert(0, ".")
from oracle import depth_port as P          # synthetic state / input generators only
from veto_b200 import depth_backbone as D
from veto_b200 import ops

B, H, W = 12, 608, 1008
out = {}
PROFILE = "--profile" in sys.argv      # under ncu: one warm-up and one step of the headline precision only
for precision in (("bf16x3",) if PROFILE else ("bf16x3", "bf16")):
    body = D.ResNetDepth(precision)
    model = torch.nn.Sequential()
    model.add_module("body", body)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in P.synth_state(0).items()})
    model = model.cuda().train()
    x = torch.from_numpy(P.synth_depth(B, H, W)).cuda()
    g = torch.randn(B, 256, *ops.depth_backbone_out_size(H, W), device="cuda")

    def step(backward):
        y = model(x)
        if backward:
            y.backward(g)

    if PROFILE:
        step(True)
        torch.cuda.synchronize()
        step(True)
        torch.cuda.synchronize()
        break
    for backward in (False, True):
        for _ in range(3):
            step(backward)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 5
        l0 = ops.launch_count()
        e0.record()
        for _ in range(n):
            step(backward)
        e1.record()
        torch.cuda.synchronize()
        out[f"{precision}_{'fwd_bwd' if backward else 'fwd'}_ms"] = round(e0.elapsed_time(e1) / n, 3)
        out[f"{precision}_{'fwd_bwd' if backward else 'fwd'}_launches"] = (ops.launch_count() - l0) // n
    with ops.StageTimer() as t:
        step(True)
    out[precision + "_stages"] = {k: round(v, 3) for k, v in t.ms.items() if v > 0}
out["batch"] = [B, 1, H, W]
print(json.dumps(out))
