"""Bring-up sweep of the MN-major descriptor geometry of gemm_tn2 (run on the GPU box; not a pytest file)."""
import itertools, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from veto_b200 import ops
torch.manual_seed(0)
dev = torch.device("cuda:0")
rows, Nw, Kw = 200, 256, 128
y = torch.randn(rows, Nw, device=dev)
x = torch.randn(rows, Kw, device=dev)
ref = y.double().t() @ x.double()
cands = [(8192, 1024, 2048), (1024, 8192, 2048), (8192, 1024, 256), (1024, 8192, 256), (8192, 2048, 2048), (16384, 1024, 2048),
         (8192, 1024, 4096), (128, 1024, 2048), (1024, 128, 2048), (8192, 128, 2048)]
for geo in cands:
    try:
        out = ops.test_gemm_tn(y, x, "bf16x3", 1, geo)
        torch.cuda.synchronize()
        err = float((out.double() - ref).abs().max() / ref.abs().max())
        print(geo, "err", f"{err:.3e}", flush=True)
    except Exception as e:
        print(geo, "EXC", str(e)[:200], flush=True)
        break
