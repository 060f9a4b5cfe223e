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

# (key prefix of the conv, key prefix of its BatchNorm, stride, padding) in the module order of include/veto_b200.h
CONVS = [("conv1", "bn1", 2, 3)]
BLOCKS = []          # (index of conv1, index of conv2, index of the downsample conv or -1)
for _l, _stride in (("layer1", 1), ("layer2", 2), ("layer3", 2)):
    for _b in (0, 1):
        p = f"{_l}.{_b}."
        s = _stride if _b == 0 else 1
        c1 = len(CONVS)
        CONVS.append((p + "conv1", p + "bn1", s, 1))
        CONVS.append((p + "conv2", p + "bn2", 1, 1))
        ds = -1
        if _b == 0 and _l != "layer1":
            ds = len(CONVS)
            CONVS.append((p + "downsample.0", p + "downsample.1", s, 0))
        BLOCKS.append((c1, c1 + 1, ds))
CHANNELS = [(1, 64, 7)] + [(64, 64, 3)] * 4 + [(64, 128, 3), (128, 128, 3), (64, 128, 1), (128, 128, 3), (128, 128, 3),
                                              (128, 256, 3), (256, 256, 3), (128, 256, 1), (256, 256, 3), (256, 256, 3)]


def out_size(height, width):
    """Spatial size of the layer3 output (four stride-2 stages, each floor((n - 1) / 2) + 1)."""
    for _ in range(4):
        height, width = (height - 1) // 2 + 1, (width - 1) // 2 + 1
    return height, width


def state_keys(prefix="body."):
    """state-dict keys in the reference's order."""
    keys = []
    order = [0] + [i for c1, c2, ds in BLOCKS for i in ((c1, c2) if ds < 0 else (c1, c2, ds))]
    for i in order:
        c, b, _, _ = CONVS[i]
        keys.append(prefix + c + ".weight")
        keys += [prefix + b + s for s in (".weight", ".bias", ".running_mean", ".running_var", ".num_batches_tracked")]
    return keys


def synth_state(seed=0, prefix="body."):
    """A deterministic (numpy) state: He-scaled convolutions, BatchNorm affine near (1, 0), non-trivial running stats."""
    rng = np.random.RandomState(seed)
    sd = {}
    for (c, b, _, _), (cin, cout, k) in zip(CONVS, CHANNELS):
        sd[prefix + c + ".weight"] = (rng.standard_normal((cout, cin, k, k)) * np.sqrt(2.0 / (k * k * cout))).astype(np.float32)
        sd[prefix + b + ".weight"] = (1.0 + 0.2 * rng.standard_normal(cout)).astype(np.float32)
        sd[prefix + b + ".bias"] = (0.1 * rng.standard_normal(cout)).astype(np.float32)
        sd[prefix + b + ".running_mean"] = (0.1 * rng.standard_normal(cout)).astype(np.float32)
        sd[prefix + b + ".running_var"] = (1.0 + 0.3 * rng.random_sample(cout)).astype(np.float32)
        sd[prefix + b + ".num_batches_tracked"] = np.array(0, np.int64)
    return sd


def synth_depth(batch, height, width, seed=0):
    """A smooth-ish synthetic depth image batch [B,1,H,W] (low-frequency ramps + noise)."""
    rng = np.random.RandomState(1000 + seed)
    yy, xx = np.meshgrid(np.linspace(0, 1, height), np.linspace(0, 1, width), indexing="ij")
    out = np.empty((batch, 1, height, width), np.float32)
    for b in range(batch):
        a = rng.standard_normal(4)
        out[b, 0] = a[0] * yy + a[1] * xx + 0.5 * np.sin(6.0 * a[2] * xx * yy) + 0.3 * rng.standard_normal((height, width))
    return out


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
