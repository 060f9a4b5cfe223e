"""CPU tests of the host-side mirror of the reference's plugin interface (no GPU needed): registries, constructor
contracts, state_dict keys, config handling, BoxList arithmetic, image sharding and the world_size-2 gloo path."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import veto_b200
from tests import harness as H
from tests.cases import CASES, load_golden
from veto_b200 import config as vcfg
from veto_b200 import distributed as vdist
from veto_b200 import registry, synth
from veto_b200.structures import BoxList

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_registry_names_and_selection():
    preds, extractors = veto_b200.load_modules()
    assert set(preds) == {"VETOPredictor", "VETOPredictor_MEET"}          # roi_relation_predictors.py:3997,3876
    assert set(extractors) == {"VETOFeatureExtractor"}                    # roi_box_feature_extractors.py:75
    cfg = vcfg.default_cfg()
    assert type(registry.make_roi_relation_predictor(cfg, 512)).__name__ == "VETOPredictor"
    cfg.MODEL.ROI_RELATION_HEAD.PREDICTOR = "VETOPredictor_MEET"
    assert type(registry.make_roi_relation_predictor(cfg, 512)).__name__ == "VETOPredictor_MEET"
    fe = registry.make_roi_box_feature_extractor(cfg, 256, for_relation=True)
    assert fe.out_channels == 256 and (fe.k_min, fe.k_max) == (2, 5) and fe.depth_scale == 0.0625
    with pytest.raises(AssertionError):                                    # utils/registry.py:4-6 semantics
        registry.ROI_RELATION_PREDICTOR.register("VETOPredictor", object)


def test_state_dict_keys_match_reference_checkpoints():
    """Keys and shapes are those of the reference modules, so reference checkpoints load strictly."""
    cfg = vcfg.default_cfg()
    p = registry.make_roi_relation_predictor(cfg, 512)
    sd = synth.predictor_state(3)
    assert set(p.state_dict()) == set(sd)
    p.load_state_dict(synth.to_torch_state(sd), strict=True)
    assert sum(v.numel() for v in p.parameters()) == 17_608_299            # SURVEY.md §6 probe: 17.608 M
    for ds, gs, n_obj in (("VG", [4, 6, 9, 19, 12], 151), ("GQA", [5, 10, 20, 65], 201)):
        cfg = H.make_cfg("VETOPredictor_MEET", dataset=ds)
        m = registry.make_roi_relation_predictor(cfg, 512)
        sd = synth.meet_state(4, n_obj, gs)
        assert set(m.state_dict()) == set(sd)
        m.load_state_dict(synth.to_torch_state(sd), strict=True)
        assert [h.out_features for h in m.model.rel_out] == [n + 2 for n in gs]
        assert m.incre_idx_list == list(load_golden("meet_gqa" if ds == "GQA" else "meet_vg")["incre_idx_list"])
    cfg = H.make_cfg("VETOPredictor_MEET")
    cfg.ENSEMBLE_LEARNING.EXPERT_GROUP = True                              # defaults.py:864: 3 experts per group
    m = registry.make_roi_relation_predictor(cfg, 512)
    sd = synth.meet_state(5, 151, [4, 6, 9, 19, 12], experts_per_group=3, expert_group=True)
    assert set(m.state_dict()) == set(sd)
    assert [k for k, _ in m.model._head_sets()][:6] == ["group_01", "group_11", "group_21", "group_31", "group_41", "group_02"]


def test_unsupported_architecture_is_refused():
    cfg = vcfg.default_cfg()
    cfg.MODEL.ROI_RELATION_HEAD.VETOTRANSFORMER.T_INPUT_DIM = 512
    with pytest.raises(RuntimeError, match="576"):
        registry.make_roi_relation_predictor(cfg, 512)
    cfg = vcfg.default_cfg()
    cfg.GLOBAL_SETTING.BETA_LOSS = True
    with pytest.raises(FileNotFoundError):           # the predicate counts must be supplied (VETO_B200.PRED_COUNTS)
        registry.make_roi_relation_predictor(cfg, 512)


def test_beta_loss_weights_match_reference():
    """GLOBAL_SETTING.BETA_LOSS (roi_relation_predictors.py:4057-4066): the CE class weights the unmodified reference
    builds from its pred_counts.pkl (golden: make_golden.py run_sample_rates) — bit-exact."""
    import os
    from tests.cases import GOLDEN_DIR
    cfg = vcfg.default_cfg()
    cfg.GLOBAL_SETTING.BETA_LOSS = True
    cfg.VETO_B200.PRED_COUNTS = os.path.join(GOLDEN_DIR, "pred_counts_vg.npy")
    m = registry.make_roi_relation_predictor(cfg, 512)
    ref = load_golden("meet_sample_rates")["beta_loss_weight"]
    assert np.array_equal(m.criterion_loss_rel.weight.numpy(), ref)
    assert "criterion_loss_rel.weight" in m.state_dict()


def test_boxlist_conventions():
    b = BoxList(torch.tensor([[10., 20., 49., 79.]]), (800, 592))
    assert b.convert("xywh").bbox.tolist() == [[10., 20., 40., 60.]]        # +1 (bounding_box.py:72-75)
    assert b.area().tolist() == [2400.]                                     # +1 (bounding_box.py:249-253)
    assert b.convert("xywh").convert("xyxy").bbox.tolist() == b.bbox.tolist()
    b.add_field("labels", torch.tensor([3]))
    assert b.has_field("labels") and b.fields() == ["labels"] and len(b) == 1
    with pytest.raises(ValueError):
        BoxList(torch.zeros(3), (1, 1))


def test_config_defaults_follow_veto_final_yaml():
    cfg = vcfg.default_cfg()
    t = cfg.MODEL.ROI_RELATION_HEAD.VETOTRANSFORMER
    assert (t.PATCH_SIZE, t.T_INPUT_DIM, t.ENC_LAYERS, t.NHEADS) == (2, 576, 6, 6)
    assert cfg.MODEL.ROI_RELATION_HEAD.POOLER_RESOLUTION == 8 and cfg.MODEL.ROI_RELATION_HEAD.MAX_PROPOSAL_PAIR == 2048
    assert vcfg.num_classes(cfg) == (151, 51)
    cfg.GLOBAL_SETTING.DATASET_CHOICE = "GQA"
    assert vcfg.num_classes(cfg) == (201, 101)
    assert vcfg.get(cfg, "VETO_B200.PRECISION") == "f16c8" and vcfg.get(cfg, "NOT.THERE", 7) == 7
    with pytest.raises(KeyError):
        cfg.merge_from_list(["MODEL.NOPE", 1])
    assert vcfg.GROUP_SPLITS[("VG", "divide4")] == synth.GROUP_SPLITS[("VG", "divide4")]


def test_image_sharding_balances_pairs():
    n_boxes = [20] * 48 + [80] * 48                                          # BASELINE.json configs[4]
    for world in (1, 2, 4, 8):
        shards = vdist.shard_images(n_boxes, world)
        assert sorted(i for s in shards for i in s) == list(range(96))
        load = [sum(n_boxes[i] * (n_boxes[i] - 1) for i in s) for s in shards]
        assert max(load) - min(load) <= 80 * 79
    assert vdist.shard_images([5, 1, 0], 2) == [[0], [1, 2]]


_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, {root!r})
from veto_b200 import distributed as vdist
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
rank = dist.get_rank()
n_boxes = [3, 9, 4, 2, 7]
mine = vdist.shard_images(n_boxes, 2)[rank]
# stand-in for the per-image results of this rank: one row per pair, tagged with (image, pair index)
rows = torch.tensor([[i, r] for i in mine for r in range(n_boxes[i] * (n_boxes[i] - 1))], dtype=torch.int64).reshape(-1, 2)
parts = vdist.gather_rows(rows)
allrows = torch.cat(parts)
assert allrows.shape[0] == sum(n * (n - 1) for n in n_boxes), allrows.shape
assert sorted(set(allrows[:, 0].tolist())) == list(range(5))
t = vdist.max_over_ranks(1.0 + rank, "cpu")
assert t == 2.0
# the training step's gradient exchange: one flat bucket, averaged over ranks (DDP semantics)
ps = [torch.nn.Parameter(torch.zeros(3, 4)), torch.nn.Parameter(torch.zeros(5)), torch.nn.Parameter(torch.zeros(2))]
ps[0].grad = torch.full((3, 4), 1.0 + rank); ps[1].grad = torch.arange(5.0) * (rank + 1)   # ps[2] has no gradient
vdist.allreduce_gradients(ps)
assert torch.equal(ps[0].grad, torch.full((3, 4), 1.5)) and torch.equal(ps[1].grad, torch.arange(5.0) * 1.5) and ps[2].grad is None
# gradients that are views of ONE flat buffer (what the relation head and the depth backbone hand to autograd): the
# buffer is reduced in place, as one bucket, padding included; the views see the averaged values
flat = torch.arange(20.0) * (rank + 1)
qs = [torch.nn.Parameter(torch.zeros(2, 3)), torch.nn.Parameter(torch.zeros(4))]
qs[0].grad = flat[2:8].view(2, 3); qs[1].grad = flat[12:16]
assert len(vdist._flat_buckets([q.grad for q in qs])) == 1
works = vdist.allreduce_gradients(qs, async_op=True)
vdist.finish_gradient_sync(works)
assert torch.equal(flat[2:16], torch.arange(20.0)[2:16] * 1.5) and torch.equal(flat[:2], torch.arange(2.0) * (rank + 1))
assert qs[0].grad.data_ptr() == flat[2:].data_ptr() and torch.equal(qs[1].grad, torch.arange(12.0, 16.0) * 1.5)
works = vdist.allreduce_flat(flat)
vdist.finish_gradient_sync(works)
dist.barrier()
dist.destroy_process_group()
print("ok", rank)
"""


def test_two_rank_gloo_sharding_and_gather(tmp_path):
    """The N>1 host path on CPU: world_size 2 over gloo (127.0.0.1), image sharding + fixed-layout result gather."""
    script = tmp_path / "worker.py"
    port = 29600 + os.getpid() % 200
    script.write_text(_WORKER.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=120)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all("ok" in o for o in outs)


def test_bench_reference_arm_contract():
    """bench.py --impl reference prints one JSON line with the contract keys (a tiny sample keeps this fast)."""
    import json
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--cpu-sample-pairs", "64"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "pairs/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]
    assert "configs[2]" in line["config"]["workload"] and line["steps"] == 1 and line["warmup"] == 0
    assert line["train"]["value"] > 0 and line["meet_gqa"]["value"] > 0 and line["sweep96"]["value"] > 0


def test_mean_recall_from_first_match_matches_reference():
    """evaluation.mean_recall (host arithmetic over the kernel's first-match ranks) against SGMeanRecall of the
    unmodified reference (sgg_eval.py:424-466), fed with the reference's own matches."""
    from tests.cases import EVAL_CASES
    from veto_b200 import evaluation as E
    c, g = EVAL_CASES["eval_recall"], load_golden("eval_recall")
    imgs = synth.make_eval_case(c["seed"], c["n_objs"], c["n_gt_rels"], c["n_pred_rels"])
    fm = [torch.from_numpy(g[f"first_match/{i}"]) for i in range(len(imgs))]
    pr = [torch.from_numpy(im["relation_tuple"][:, 2]) for im in imgs]
    mr = E.mean_recall(fm, pr, 51)
    for k in (20, 50, 100):
        assert np.allclose(mr["mean_recall_list"][k], g[f"mean_recall_list/{k}"], rtol=1e-12, atol=0)
        assert abs(mr["mean_recall"][k] - float(g[f"mean_recall/{k}"])) <= 1e-12


def test_meet_group_sampling_invariants():
    """Properties of the host-side MEET group sampling (meet_sampling.py) for random label vectors and every padding
    mode: a pair joins a PREFIX of the group heads (heads 0..a-1), at least the heads up to one before its own group;
    rows are listed once per head in ascending order; the label table is -1 exactly off the chosen rows, 0 for
    background, 1..n_k for members and n_k + 1 for foreign predicates."""
    import random
    from hypothesis import given, settings, strategies as st
    from veto_b200 import meet_sampling as MS
    from veto_b200.predictor import incre_idx_list

    sizes = vcfg.GROUP_SPLITS[("VG", "divide4")]
    incre = incre_idx_list(sizes, 51)
    rates = MS.sample_rate_matrix("VG", sizes)
    assert rates.shape == (5, 51) and np.all(rates > 0) and np.all(rates <= 1)

    @settings(max_examples=60, deadline=None)
    @given(st.lists(st.integers(0, 50), min_size=1, max_size=200), st.sampled_from(["rand_insert", "rand_choose", "all_include"]),
           st.integers(0, 2 ** 31))
    def check(labels, mode, seed):
        rng = random.Random(seed)
        chosen = MS.group_sampling(labels, incre, rates, len(sizes), mode, rng)
        member = np.zeros((len(sizes), len(labels)), bool)
        for k, rows in enumerate(chosen):
            assert rows == sorted(set(rows))
            member[k, rows] = True
        for i, p in enumerate(labels):
            col = member[:, i]
            if p == 0:
                assert col.sum() == (1 if mode == "rand_insert" else (0 if (mode == "rand_choose" and not col.any()) else len(sizes)))
            else:
                a = int(col.sum())
                assert col[:a].all() and not col[a:].any()            # a prefix of the heads
                assert a >= incre[p] - 1                              # at least every head before its own group's
        table = MS.group_local_labels(labels, chosen, incre)
        assert table.shape == (len(sizes), len(labels)) and np.array_equal(table >= 0, member)
        for k in range(len(sizes)):
            for i, p in enumerate(labels):
                if member[k, i]:
                    want = 0 if p == 0 else (p - sum(sizes[:k]) if incre[p] == k + 1 else sizes[k] + 1)
                    assert table[k, i] == want

    check()


def test_image_sharding_is_a_partition():
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=100, deadline=None)
    @given(st.lists(st.integers(0, 120), min_size=0, max_size=64), st.integers(1, 8), st.sampled_from([64, 2048, 1 << 62]))
    def check(n_boxes, world, cap):
        shards = vdist.shard_images(n_boxes, world, cap)
        assert len(shards) == world
        assert sorted(i for s in shards for i in s) == list(range(len(n_boxes)))
        assert all(s == sorted(s) for s in shards)
        loads = [sum(vdist.pair_count(n_boxes[i], cap) for i in s) for s in shards]
        if n_boxes:
            biggest = max(vdist.pair_count(n, cap) for n in n_boxes)
            assert max(loads) <= sum(loads) / world + biggest           # the LPT bound

    check()
