"""Helpers shared by the parity tests."""
import numpy as np


def tie_groups_equal(rows_a, rows_b, keys):
    """rows_* [R,k] integer rows ordered by `keys` (descending).  torch.sort in the reference is
    unstable (sampling.py:43, inference.py:443), so rows that tie on the key may come in any order:
    require the same multiset of rows inside every run of equal keys."""
    rows_a, rows_b, keys = np.asarray(rows_a), np.asarray(rows_b), np.asarray(keys)
    if rows_a.shape != rows_b.shape:
        return False
    start = 0
    for end in range(1, len(keys) + 1):
        if end == len(keys) or keys[end] != keys[start]:
            a = sorted(map(tuple, rows_a[start:end].reshape(end - start, -1)))
            b = sorted(map(tuple, rows_b[start:end].reshape(end - start, -1)))
            if a != b:
                return False
            start = end
    return True


def rel_err(a, ref):
    a, ref = np.asarray(a, np.float64), np.asarray(ref, np.float64)
    return float(np.abs(a - ref).max() / max(np.abs(ref).max(), 1e-30))
