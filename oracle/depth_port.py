"""ORACLE (test infrastructure, never on the product path) — CPU restatement of the depth backbone.

ResNetDepth (pysgg/modeling/backbone/resnet_depth.py:11-47) is torchvision's ResNet-18 (BasicBlock, [2,2,2,2]) with a
one-channel conv1 and layer4 / avgpool / fc removed; backbone.py:83-93 wraps it as nn.Sequential(body=...).  This file
restates its forward with torch.nn.functional on the state dict (keys ``body.*``) so that the training-mode gradients
come from torch autograd on the CPU in fp32 (or fp64 with ``dtype=torch.float64``).  Pinned by tests/golden/depth_backbone.npz,
which the reference's own ResNetDepth produced (tests/golden/make_golden.py run_depth_backbone).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

# the module table and the seeded numpy generators live with the other synthetic inputs (veto_b200/synth.py) so that the
# product bench needs nothing from oracle/; re-exported here for the tests
from veto_b200.synth import (DEPTH_BLOCKS as BLOCKS, DEPTH_CHANNELS as CHANNELS, DEPTH_CONVS as CONVS,  # noqa: E402,F401
                             depth_out_size as out_size, depth_state as synth_state, depth_state_keys as state_keys,
                             depth_images as synth_depth)


def forward(state, depth, training, prefix="body.", momentum=0.1, eps=1e-5):
    """ResNetDepth.forward (resnet_depth.py:36-47; BasicBlock.forward of torchvision/models/resnet.py).  ``state`` maps
    keys to torch tensors (leaf tensors with requires_grad for a gradient run); running statistics are updated in
    place in training mode.  Returns the layer3 output [B,256,H/16,W/16]."""

    def conv_bn(i, x, relu):
        c, b, stride, pad = CONVS[i]
        x = F.conv2d(x, state[prefix + c + ".weight"], None, stride, pad)
        x = F.batch_norm(x, state[prefix + b + ".running_mean"], state[prefix + b + ".running_var"],
                         state[prefix + b + ".weight"], state[prefix + b + ".bias"], training, momentum, eps)
        return F.relu(x) if relu else x

    x = conv_bn(0, depth, True)
    x = F.max_pool2d(x, 3, 2, 1)
    for c1, c2, ds in BLOCKS:
        identity = x if ds < 0 else conv_bn(ds, x, False)
        out = conv_bn(c2, conv_bn(c1, x, True), False)
        x = F.relu(out + identity)
    return x


def train_step(state_np, depth_np, grad_out_np, dtype=torch.float32, prefix="body."):
    """One training-mode forward + backward from a given output gradient: (output, {param key: gradient},
    {running-stat key: updated value}), numpy."""
    st = {}
    for k, v in state_np.items():
        t = torch.from_numpy(np.asarray(v))
        if t.is_floating_point():
            t = t.to(dtype).clone()
            if not ("running_" in k):
                t.requires_grad_(True)
        st[k] = t
    out = forward(st, torch.from_numpy(depth_np).to(dtype), True, prefix)
    out.backward(torch.from_numpy(grad_out_np).to(dtype))
    grads = {k: t.grad.numpy() for k, t in st.items() if t.is_floating_point() and t.requires_grad}
    stats = {k: t.detach().numpy() for k, t in st.items() if "running_" in k}
    return out.detach().numpy(), grads, stats


def eval_forward(state_np, depth_np, dtype=torch.float32, prefix="body."):
    st = {k: (torch.from_numpy(np.asarray(v)).to(dtype) if np.asarray(v).dtype.kind == "f" else torch.from_numpy(np.asarray(v)))
          for k, v in state_np.items()}
    with torch.no_grad():
        return forward(st, torch.from_numpy(depth_np).to(dtype), False, prefix).numpy()
