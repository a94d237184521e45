"""GPU parity: federated aggregation + fused SGD (fp32 tolerance 1e-6) and the fairness-metric kernel
(integer counts BIT-EXACT vs the oracle; derived scores equal to the reference goldens to 1e-12)."""
from __future__ import annotations

import copy
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import ref_port as rp
from tests.golden import recipes

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"
DEV = "cuda:0"


@pytest.mark.parametrize("name", list(recipes.FEDAVG_CASES))
def test_average_weights_ema_matches_reference_golden(name):
    from fairfedmed_b200 import fed_utils
    rc = recipes.FEDAVG_CASES[name]
    gold = np.load(GOLD / "fedavg.npz")
    w_g, w_loc, n_k, n_kg = recipes.fedavg_inputs(rc)
    to = lambda d: {k: v.to(DEV) for k, v in d.items()}
    out = fed_utils.average_weights_EMA(to(w_g), [to(w) for w in w_loc], rc["idxs"], n_k, n_kg, rc["epoch"],
                                        rc["max_epoch"], shared_half_s=rc["shared_half_s"])
    assert set(out) == set(w_g)
    for k, v in out.items():
        np.testing.assert_allclose(v.cpu().numpy(), gold[f"{name}.{k}"], rtol=1e-6, atol=1e-7)
    if rc.get("plain"):
        out2 = fed_utils.average_weights([to(w) for w in w_loc], rc["idxs"], n_k, n_kg)
        for k, v in out2.items():
            np.testing.assert_allclose(v.cpu().numpy(), gold[f"{name}.plain.{k}"], rtol=1e-6, atol=1e-7)


def test_aggregation_vit_b16_size_idempotent_and_convex():
    """P = 1 110 880 (ViT-B/16 adapters): averaging identical clients returns them (weights sum to one) and the
    EMA with beta_decay = 0 ignores the previous global."""
    from fairfedmed_b200 import fed_utils
    G, r = 3, 12
    g = torch.Generator().manual_seed(0)
    sd = {"prompt_learner.ctx": torch.randn(2, 4, 512, generator=g)}
    for i in range(12):
        for l, (cin, cout) in (("c_fc", (768, 3072)), ("c_proj", (3072, 768))):
            p = f"image_encoder.transformer.resblocks.{i}.mlp.{l}."
            sd[p + "lora_A.weight"] = torch.randn(cin, r, generator=g)
            sd[p + "lora_S.weight"] = torch.rand(G, r, generator=g)
            sd[p + "lora_B.weight"] = torch.randn(r, cout, generator=g)
    assert sum(v.numel() for v in sd.values()) == 1_110_880
    sd = {k: v.to(DEV) for k, v in sd.items()}
    n_kg = [[5, 7, 9], [11, 2, 4], [3, 3, 3]]
    out = fed_utils.average_weights_EMA({k: torch.zeros_like(v) for k, v in sd.items()}, [sd, sd, sd], [0, 1, 2],
                                        [21, 17, 9], n_kg, 0, 50, shared_half_s=False)
    for k in sd:
        torch.testing.assert_close(out[k], sd[k], rtol=2e-6, atol=1e-6)


def test_fused_sgd_double_step_matches_torch():
    from fairfedmed_b200 import ops
    g = torch.Generator().manual_seed(2)
    n = 100_003
    w0 = torch.randn(n, generator=g)
    w_ref = w0.clone().requires_grad_(True)
    opt = torch.optim.SGD([w_ref], lr=1e-3, momentum=0.9, weight_decay=5e-4)
    w = w0.clone().to(DEV)
    mom = torch.zeros(n, device=DEV)
    first = True
    for step in range(3):
        grad = torch.randn(n, generator=g)
        w_ref.grad = grad.clone()
        opt.step()
        opt.step()                                    # the reference steps the shared optimizer twice (F6)
        ops.sgd_step_(w, grad.to(DEV), mom, 1e-3, 0.9, 5e-4, 2, first)
        first = False
        torch.testing.assert_close(w.cpu(), w_ref.detach(), rtol=1e-6, atol=1e-7)


def _oracle_counts(prob, y, attrs, max_groups):
    n_attr = attrs.shape[0]
    table = np.zeros((1 + n_attr * (max_groups + 1), 8), dtype=np.int64)

    def fill(slot, mask):
        if mask.sum() == 0:
            return
        p, t = prob[mask], y[mask]
        g0, e0, _, _ = rp.mann_whitney_counts(p[:, 0], t == 0)
        g1, e1, _, _ = rp.mann_whitney_counts(p[:, 1], t == 1)
        table[slot, :4] = (g0, e0, g1, e1)
        table[slot, 4:] = rp.confusion_counts(p, t)

    fill(0, np.ones_like(y, dtype=bool))
    for a in range(n_attr):
        for gidx in range(-1, max_groups):
            fill(1 + a * (max_groups + 1) + gidx + 1, attrs[a] == gidx)
    return table


@pytest.mark.parametrize("name", list(recipes.METRIC_CASES))
def test_group_counts_bit_exact_and_scores_match_reference(name):
    from fairfedmed_b200 import metrics as M
    rc = recipes.METRIC_CASES[name]
    gold = np.load(GOLD / "metrics.npz")
    prob, y, attrs = recipes.metric_inputs(rc)
    mg = max(rc["groups"])
    counts = M.group_counts(prob, y, attrs, max_groups=mg)
    np.testing.assert_array_equal(counts.t, _oracle_counts(prob, y, attrs, mg))        # integers: exact
    assert M.compute_auc(prob, y) == pytest.approx(float(gold[f"{name}.auc"]), abs=1e-12)
    assert M.compute_auc(prob[:, 1], y) == pytest.approx(float(gold[f"{name}.auc_binary"]), abs=1e-12)
    res = M.evalute_comprehensive_perf_scores(prob, y, attrs)
    assert res[0] == pytest.approx(float(gold[f"{name}.overall_acc"]), abs=1e-12)
    np.testing.assert_allclose(res[1], gold[f"{name}.esaccs"], atol=1e-12)
    assert res[2] == pytest.approx(float(gold[f"{name}.overall_auc"]), abs=1e-12)
    np.testing.assert_allclose(res[3], gold[f"{name}.esaucs"], atol=1e-12)
    np.testing.assert_allclose(res[8], gold[f"{name}.disparity"], atol=1e-12)
    for a in range(attrs.shape[0]):
        np.testing.assert_allclose(res[4][a], gold[f"{name}.gauc{a}"], atol=1e-12)
        assert M.equity_scaled_AUC(prob, y, attrs[a]) == pytest.approx(float(gold[f"{name}.esauc{a}"]), abs=1e-12)
        assert M.equity_scaled_accuracy(prob, y, attrs[a]) == pytest.approx(float(gold[f"{name}.esacc{a}"]), abs=1e-12)
    # DPD / EOD / AOD: "parity unpinned" upstream; must equal the oracle's restatement of the public definitions
    ref = rp.comprehensive_scores(prob, y, attrs)
    np.testing.assert_allclose(res[5], ref[5], atol=1e-12)
    np.testing.assert_allclose(res[6], ref[6], atol=1e-12)
    np.testing.assert_allclose(res[7], ref[7], atol=1e-12)


def test_group_counts_large_and_degenerate():
    from fairfedmed_b200 import metrics as M
    rng = np.random.default_rng(0)
    N = 200_000
    y = rng.integers(0, 2, N)
    p1 = np.round(rng.random(N), 3).astype(np.float32)            # ~1000 distinct scores: long tie runs
    prob = np.stack([1 - p1, p1], axis=1).astype(np.float32)
    attrs = rng.integers(-1, 3, (2, N))
    counts = M.group_counts(prob, y, attrs, max_groups=3)
    np.testing.assert_array_equal(counts.t, _oracle_counts(prob, y, attrs, 3))
    # checksum-of-checksums: group confusion counts (incl. unknown) add up to the overall slot
    for a in range(2):
        tot = sum(counts.slot(a, g)[4:] for g in range(-1, 3))
        np.testing.assert_array_equal(tot, counts.slot()[4:])
    # all scores equal -> AUC exactly 0.5; a single-class group has no AUC
    flat = np.full((64, 2), 0.5, dtype=np.float32)
    yy = np.arange(64) % 2
    assert M.compute_auc(flat, yy) == 0.5
    with pytest.raises(ValueError):
        M.compute_auc(flat, np.zeros(64, dtype=np.int64))
    one = M.group_counts(flat[:1], yy[:1])                       # N = 1
    assert int(one.slot()[4:].sum()) == 1
