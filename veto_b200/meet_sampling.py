"""Training-time group sampling of VETOPredictor_MEET.

The reference decides, pair by pair in a Python loop with one ``.item()`` device sync per pair
(roi_relation_predictors.py:3940-3969), which group heads a training pair contributes to, then relabels the chosen
pairs per group with another per-element loop (:3812-3821).  The product path does both on the device
(``veto_meet_group_labels``, csrc/meet_sample.cu, via ``ops.meet_group_labels``): a dense ``[n_groups, R]`` table of
group-local labels (-1 = pair not in that group's loss) that ``veto_relation_train_step`` consumes, nothing crossing to
the host.  This module builds the small tables that kernel reads (``sample_rate_matrix``, ``local_label_table``) and
keeps the host restatement of the two loops (``group_sampling``, ``group_local_labels``) — the checker the tests compare
the kernel with, and the way to replay the reference's seeded ``random`` stream (``reference_draws``).

Tables: ``PREDICATE_COUNTS`` are the per-predicate training-set frequencies the reference hard-codes
(SHA_GCL_extra/extra_function_utils.py:186-205, predicates in the frequency-sorted order the MEET datasets use).
"""
from __future__ import annotations

import random as _random
from typing import List, Sequence

import numpy as np

PREDICATE_COUNTS = {
    "VG": [3024465, 109355, 67144, 47326, 31347, 21748, 15300, 10011, 11059, 10764, 6712, 5086, 4810, 3757, 4260, 3167,
           2273, 1829, 1603, 1413, 1225, 793, 809, 676, 352, 663, 752, 565, 504, 644, 601, 551, 460, 394, 379, 397, 429,
           364, 333, 299, 270, 234, 171, 208, 163, 157, 151, 71, 114, 44, 4],
    "GQA": [200000, 64218, 47205, 32126, 25203, 21104, 15890, 15676, 7688, 6966, 6596, 6044, 5250, 4260, 4180, 4131,
            2859, 2559, 2368, 2351, 2134, 1673, 1532, 1373, 1273, 1175, 1139, 1123, 1077, 941, 916, 849, 835, 808, 782,
            767, 628, 603, 569, 540, 494, 416, 412, 412, 398, 395, 394, 390, 345, 327, 302, 301, 292, 275, 270, 267,
            267, 264, 258, 251, 233, 233, 229, 224, 215, 214, 209, 204, 198, 195, 192, 191, 185, 181, 176, 158, 158,
            154, 151, 148, 143, 136, 131, 130, 130, 128, 127, 125, 124, 124, 121, 118, 112, 112, 106, 105, 104, 103,
            102, 52, 52],
}


def sample_rate_matrix(dataset: str, group_sizes: Sequence[int]) -> np.ndarray:
    """generate_sample_rate_vector_sep2 (SHA_GCL_extra/extra_function_utils.py:185-257) as a float64 [n_groups,
    num_rel] table: row g = the probability with which a pair of predicate p is kept for the group set {0..g}.

    For group g (predicates lo+1 .. hi in frequency order) with median count m over its own predicates:
      * a predicate of the group or of an earlier group is kept with min(1, max(0.01, m / count)) — frequent
        predicates are down-sampled towards the group's median;
      * the background class with max(0.01, 10 m / count[0]);
      * predicates of later groups with 1 (their count never exceeds the background's, :240-251)."""
    counts = np.asarray(PREDICATE_COUNTS[dataset], dtype=np.float64)
    bounds = np.concatenate([[0], np.cumsum(group_sizes)]).astype(int)
    out = np.zeros((len(group_sizes), counts.size), dtype=np.float64)
    for g in range(len(group_sizes)):
        lo, hi = bounds[g], bounds[g + 1]
        median = float(np.median(counts[lo + 1:hi + 1]))
        seen = counts[1:hi + 1]
        out[g, 1:hi + 1] = np.where(seen > median, np.maximum(median / seen, 0.01), 1.0)
        if counts[0] > median:
            out[g, 0] = max(median / counts[0] * 10.0, 0.01)   # the reference's operation order (bit-exact table)
        else:                                   # never true for the real tables; the reference's indexing quirk (:222-224)
            out[g, 0] = 1.0 if lo == 0 else 0.0
        ceiling = max(counts[0], counts[lo + 1:hi + 1].max())
        later = counts[hi + 1:]
        rate = np.maximum(median / later, 0.01)
        if later.size:
            rate[0] = max(median / later[0] * 10.0, 0.01)
        out[g, hi + 1:] = np.where(later > ceiling, rate, 1.0)
    return out


def group_sampling(rel_labels: Sequence[int], incre_idx: Sequence[int], rates: np.ndarray, n_groups: int,
                   zero_mode: str = "rand_insert", rng=_random) -> List[List[int]]:
    """cur_chosen_matrix of VETOPredictor_MEET.forward (roi_relation_predictors.py:3940-3969): chosen[k] = the pair
    rows whose loss head k sees, in ascending row order.  `rng` must offer random() and randint(a, b) — the
    reference uses the global ``random`` module, and one draw per pair in row order is kept so that seeding
    ``random`` reproduces the reference's choice.

    Foreground pair of predicate p (group g_p = incre_idx[p], 1-based): with one uniform u, walk a = n_groups .. 1
    and stop at the first a with u <= rates[a-1][p] or a < g_p; the pair then joins heads 0 .. a-1.
    Background pair: 'rand_insert' = one uniformly drawn head; 'rand_choose' = all heads with probability 0.6;
    'all_include' = all heads."""
    chosen: List[List[int]] = [[] for _ in range(n_groups)]
    for i, p in enumerate(rel_labels):
        if p == 0:
            if zero_mode == "rand_insert":
                chosen[rng.randint(0, n_groups - 1)].append(i)
            elif zero_mode in ("rand_choose", "all_include"):
                if zero_mode == "all_include" or rng.random() >= 0.4:
                    for k in range(n_groups):
                        chosen[k].append(i)
            continue
        g_p = incre_idx[p]
        u = rng.random()
        for a in range(n_groups, 0, -1):
            if u <= rates[a - 1][p] or a < g_p:
                for k in range(a):
                    chosen[k].append(i)
                break
    return chosen


def group_local_labels(rel_labels: Sequence[int], chosen: Sequence[Sequence[int]], incre_idx: Sequence[int]) -> np.ndarray:
    """Ensemble.forward's relabelling (roi_relation_predictors.py:3806-3821) as an int64 [n_groups, R] table:
    for head k with member predicates P_k (ascending ids): 0 stays 0, a member becomes its 1-based position in P_k,
    any other foreground predicate becomes len(P_k) + 1 (the head's out-of-group class); rows not chosen get -1."""
    labels = np.asarray(rel_labels, dtype=np.int64)
    incre = np.asarray(incre_idx, dtype=np.int64)
    out = np.full((len(chosen), labels.size), -1, dtype=np.int64)
    for k, rows in enumerate(chosen):
        members = np.nonzero(incre == k + 1)[0]
        position = np.full(incre.size, len(members) + 1, dtype=np.int64)
        position[members] = np.arange(1, len(members) + 1)
        position[0] = 0
        rows = np.asarray(rows, dtype=np.int64)
        out[k, rows] = position[labels[rows]]
    return out


def local_label_table(incre_idx: Sequence[int], n_groups: int) -> np.ndarray:
    """int32 [n_groups, num_rel]: head k's local label of global predicate p (roi_relation_predictors.py:3806-3821):
    0 for the background, the 1-based position of p among the head's member predicates, len(members) + 1 otherwise."""
    incre = np.asarray(incre_idx, dtype=np.int64)
    out = np.zeros((n_groups, incre.size), dtype=np.int32)
    for k in range(n_groups):
        members = np.nonzero(incre == k + 1)[0]
        out[k, :] = len(members) + 1
        out[k, members] = np.arange(1, len(members) + 1)
        out[k, 0] = 0
    return out


def reference_draws(rel_labels: Sequence[int], n_groups: int, zero_mode: str = "rand_insert", rng=_random):
    """The draws the reference's loop (:3940-3969) takes from ``random`` for these labels, in its order: (u float64 [R],
    bg_head int32 [R]) to inject into ``ops.meet_group_labels`` so that a run seeded like the reference picks the same
    pairs.  One draw per pair: randint for a 'rand_insert' background pair, random() for a foreground pair and for a
    'rand_choose' background pair, none for 'all_include' backgrounds."""
    u = np.zeros(len(rel_labels), dtype=np.float64)
    head = np.zeros(len(rel_labels), dtype=np.int32)
    for i, p in enumerate(rel_labels):
        if p == 0:
            if zero_mode == "rand_insert":
                head[i] = rng.randint(0, n_groups - 1)
            elif zero_mode == "rand_choose":
                u[i] = rng.random()
        else:
            u[i] = rng.random()
    return u, head


class LazyExpertDist:
    """What VETOPredictor_MEET.forward returns as its 5th value in train() mode: the reference hands back
    ``expert_dist`` = the chosen-rows lists, appended once per pair (:3969).  Nothing on the training path reads it
    (relation_head.py:196-203), so the device table is converted to those lists only if somebody indexes it."""

    def __init__(self, table):
        self.table = table            # int64 [n_groups, R] on the device
        self._lists = None

    def chosen(self):
        if self._lists is None:
            t = self.table.cpu().numpy()
            self._lists = [np.nonzero(row >= 0)[0].tolist() for row in t]
        return self._lists

    def __len__(self):
        return int(self.table.shape[1])

    def __getitem__(self, i):
        if not -len(self) <= i < len(self):
            raise IndexError(i)
        return self.chosen()

    def __iter__(self):
        return (self.chosen() for _ in range(len(self)))
