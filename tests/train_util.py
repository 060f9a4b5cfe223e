"""Helpers shared by the CPU oracle test and the GPU parity tests of the training step."""
from __future__ import annotations

import numpy as np

from tests.cases import case_batch, case_class_weight, case_rel_labels, case_state


def train_case_inputs(c):
    """(batch, state_dict numpy incl. the CE class weight, per-image pair arrays, per-image rel_labels)."""
    from oracle import veto_oracle as O
    batch = case_batch(c)
    sd = dict(case_state(c))
    cw = case_class_weight(c)
    if cw is not None:
        sd["criterion_loss_rel.weight"] = cw
    pairs = O.prepare_test_pairs(batch["n_boxes"])
    labels = case_rel_labels(c, [len(p) for p in pairs])
    return batch, sd, pairs, labels


def oracle_train_step(c, batch, sd, pairs, rel_labels, drop=None):
    """oracle/torch_port.train_step on a case: dict(loss, grads{key: np}, g_roi_depth, bn running stats, logits)."""
    import torch
    from oracle import torch_port as TP
    tsd = TP.to_torch(sd)
    boxes = [torch.from_numpy(b) for b in batch["boxes"]]
    x2d, d2d = TP.pooler_forward([torch.from_numpy(f) for f in batch["feats"]], torch.from_numpy(batch["depth"]), boxes)
    loss, grads, g_d2d, g_x2d, bn_out, logits = TP.train_step(
        tsd, boxes, [torch.from_numpy(p) for p in pairs], [torch.from_numpy(l) for l in rel_labels], x2d, d2d, c["mode"],
        labels=[torch.from_numpy(l) for l in batch["labels"]],
        predict_logits=[torch.from_numpy(l) for l in batch.get("predict_logits", [])] or None,
        class_weight=tsd["criterion_loss_rel.weight"], drop=drop)
    return dict(loss=float(loss), grads={k: v.numpy() for k, v in grads.items()}, g_roi_depth=g_d2d.numpy(),
                g_roi_rgb=g_x2d.numpy(), running_mean=bn_out["running_mean"].numpy(),
                running_var=bn_out["running_var"].numpy(), logits=logits.numpy(), x2d=x2d.numpy(), d2d=d2d.numpy())


def grad_error(mine: np.ndarray, ref: np.ndarray) -> float:
    """max |mine - ref| relative to max |ref| (the logit metric of north_star, applied to a gradient tensor)."""
    scale = max(float(np.abs(ref).max()), 1e-30)
    return float(np.abs(mine.astype(np.float64) - ref.astype(np.float64)).max() / scale)


def check_against_golden(grads, g, tol, skip=()):
    """Compare full gradient tensors with the norm / sum / samples the fixture keeps of the reference's gradients."""
    worst = {}
    for key in g.files:
        if not key.startswith("gstat/"):
            continue
        k = key[6:]
        if k in skip or k not in grads:
            continue
        mine = grads[k].reshape(-1).astype(np.float64)
        norm_ref, sum_ref = g[key]
        idx, val = g["gidx/" + k], g["gval/" + k].astype(np.float64)
        scale = max(np.abs(val).max(), norm_ref / np.sqrt(mine.size), 1e-30)
        e_samples = np.abs(mine[idx] - val).max() / scale
        e_norm = abs(np.sqrt((mine ** 2).sum()) - norm_ref) / max(norm_ref, 1e-30)
        worst[k] = max(e_samples, e_norm)
        assert e_samples <= tol and e_norm <= tol, f"{k}: samples {e_samples:.2e} norm {e_norm:.2e} > {tol}"
    return worst


def meet_train_case_inputs(c):
    """(batch, MEET state_dict numpy, per-image pair arrays, concatenated rel_labels list, obj_preds) of a MEET case."""
    from oracle import veto_oracle as O
    batch = case_batch(c)
    sd = dict(case_state(c))
    pairs = O.prepare_test_pairs(batch["n_boxes"])
    labels = case_rel_labels(c, [len(p) for p in pairs])
    return batch, sd, pairs, labels


def oracle_meet_train_step(c, batch, sd, pairs, head_labels, drop=None):
    """oracle/torch_port.train_step_meet on a MEET case: dict(losses [G], grads{key: np}, g_roi_depth)."""
    import torch
    from oracle import torch_port as TP
    tsd = TP.to_torch(sd)
    boxes = [torch.from_numpy(b) for b in batch["boxes"]]
    x2d, d2d = TP.pooler_forward([torch.from_numpy(f) for f in batch["feats"]], torch.from_numpy(batch["depth"]), boxes)
    obj_preds = torch.from_numpy(np.concatenate(batch["labels"])).long()
    losses, grads, g_d2d, g_x2d, bn_out = TP.train_step_meet(
        tsd, boxes, [torch.from_numpy(p) for p in pairs], head_labels, x2d, d2d, obj_preds, drop=drop)
    return dict(losses=np.array([float(l) for l in losses]), grads={k: v.numpy() for k, v in grads.items()},
                g_roi_depth=g_d2d.numpy(), g_roi_rgb=g_x2d.numpy(), x2d=x2d.numpy(), d2d=d2d.numpy())


def check_detect_sample(cand, pairs, labels, batch_size, num_pos, per_gt=4):
    """Structural check of one image's detect_relsample output (pairs [R,2], labels [R]) against the oracle's candidate
    sets (oracle.detect_relsample_candidates): sizes as sampling.py:256-293 prescribes, every foreground row a candidate
    of a ground-truth relation with that label, all candidates present where nothing had to be drawn, background rows
    distinct members of the quality-ranked pool.  Returns (n_fg, n_bg_rows)."""
    pairs = np.asarray(pairs).reshape(-1, 2)
    labels = np.asarray(labels)
    kept = [min(len(c[3]), per_gt) for c in cand["gt"]]
    n_fg_all = sum(kept)
    n_fg = min(n_fg_all, num_pos)
    n_bg_pool = len(cand["bg"])
    num_neg = max(0, min(batch_size - n_fg, n_bg_pool))
    if n_fg + num_neg == 0:                                   # :298-304 placeholder rows
        assert pairs.tolist() == [[0, 0], [0, 0]] and labels.tolist() == [0, 0]
        return 0, 2
    assert len(pairs) == n_fg + num_neg, (len(pairs), n_fg, num_neg)
    assert np.all(labels[:n_fg] != 0) and np.all(labels[n_fg:] == 0)
    by_label = {}
    for h, t, l, c in cand["gt"]:
        by_label.setdefault(l, set()).update(c)
    fg_rows = [(int(a), int(b), int(l)) for (a, b), l in zip(pairs[:n_fg], labels[:n_fg])]
    for a, b, l in fg_rows:
        assert (a, b) in by_label.get(l, ()), (a, b, l)
    if n_fg_all <= num_pos:
        # nothing cut at the image level: relations with <= per_gt candidates contribute all of them
        from collections import Counter
        have = Counter(fg_rows)
        for h, t, l, c in cand["gt"]:
            if len(c) <= per_gt:
                for a, b in c:
                    assert have[(a, b, l)] >= 1, (h, t, l, a, b)
    bg_rows = [(int(a), int(b)) for a, b in pairs[n_fg:]]
    assert len(set(bg_rows)) == len(bg_rows)
    assert set(bg_rows) <= set(cand["bg"][: int(num_neg * 2.0)])
    return n_fg, num_neg
