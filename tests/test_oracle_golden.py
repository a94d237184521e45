"""Pin the restated oracle (oracle/ref_port.py) to outputs of the REAL reference.

The fixtures under tests/golden/*.npz were produced by tests/golden/make_golden.py, which imports the
reference from /root/reference; these tests need only the fixtures, so they also run on the GPU box.
"""
from __future__ import annotations

from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import ref_port as rp
from tests.golden import recipes

GOLD = Path(__file__).resolve().parent / "golden"


def load(name):
    return np.load(GOLD / name, allow_pickle=False)


def close(a, b, rtol=1e-5, atol=1e-6):
    a = a.detach().numpy() if torch.is_tensor(a) else np.asarray(a)
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol)


# ----------------------------------------------------------------------------------------------- FairLoRA linear
@pytest.mark.parametrize("name", list(recipes.FAIRLORA_CASES))
def test_fairlora_linear_matches_reference(name):
    rc = recipes.FAIRLORA_CASES[name]
    gold = load("fairlora.npz")
    t = recipes.fairlora_inputs(rc)
    scaling = rc["alpha"] / rc["rank"]
    leaves = {k: t[k].clone().requires_grad_(True) for k in ("x", "A", "B")}
    if rc["kind"] != "LoRA":
        leaves["S"] = t["S"].clone().requires_grad_(True)
    sg = None
    if rc["global_s"]:
        sg = t["S_global"].clone().requires_grad_(True)
    if rc["kind"] == "FairLoRA":
        y = rp.fairlora_linear(leaves["x"], t["W"], t["bias"], leaves["A"], leaves["S"], leaves["B"], t["attr"],
                               scaling, sg)
    elif rc["kind"] == "SVLoRA":
        y = rp.svlora_linear(leaves["x"], t["W"], t["bias"], leaves["A"], leaves["S"], leaves["B"], scaling, sg)
    else:
        y = rp.lora_linear(leaves["x"], t["W"], t["bias"], leaves["A"], leaves["B"], scaling)
    (y * t["dy"]).sum().backward()
    close(y, gold[f"{name}.y"], 1e-5, 1e-5)
    close(leaves["x"].grad, gold[f"{name}.dx"], 1e-5, 1e-5)
    close(leaves["A"].grad, gold[f"{name}.dA"], 1e-4, 1e-4)
    close(leaves["B"].grad, gold[f"{name}.dB"], 1e-4, 1e-4)
    if rc["kind"] != "LoRA":
        close(leaves["S"].grad.reshape(gold[f"{name}.dS"].shape), gold[f"{name}.dS"], 1e-4, 1e-4)
    if sg is not None and rc["kind"] == "FairLoRA":
        close(sg.grad, gold[f"{name}.dS_global"], 1e-4, 1e-4)
    if rc.get("merged"):
        w = rp.fairlora_merged_weight(t["W"], t["A"], t["S"], t["B"], t["attr"], scaling, rc["Bp"])
        close(w, gold[f"{name}.merged_w"], 1e-5, 1e-6)


def test_fairlora_init_matches_reference():
    gold = load("fairlora.npz")
    for name in ("vit_small", "oct_slices"):
        rc = recipes.FAIRLORA_CASES[name]
        close(rp.fairlora_init_S(rc["groups"], rc["rank"]), gold[f"{name}.init_S"], 1e-6, 1e-7)


# ----------------------------------------------------------------------------------------------- Sinkhorn / COT
@pytest.mark.parametrize("name", list(recipes.SINKHORN_CASES))
def test_sinkhorn_matches_reference(name):
    rc = recipes.SINKHORN_CASES[name]
    gold = load("sinkhorn.npz")
    K, u, v = recipes.sinkhorn_inputs(rc)
    if rc["mode"] == "Sinkhorn":
        T, iters = rp.sinkhorn(K, u, v, rc["thresh"], rc["max_iter"])
    else:
        T, iters = rp.entropic_cot(u, v, K, rc["thresh"], rc["max_iter"])
    assert 1 <= iters <= rc["max_iter"]
    close(T, gold[f"{name}.T"], 1e-5, 1e-9)


# ----------------------------------------------------------------------------------------------- aggregation
@pytest.mark.parametrize("name", list(recipes.FEDAVG_CASES))
def test_fedavg_matches_reference(name):
    rc = recipes.FEDAVG_CASES[name]
    gold = load("fedavg.npz")
    w_g, w_loc, n_k, n_kg = recipes.fedavg_inputs(rc)
    out = rp.average_weights_ema(w_g, w_loc, rc["idxs"], n_k, n_kg, rc["epoch"], rc["max_epoch"],
                                 shared_half_s=rc["shared_half_s"])
    for k, v in out.items():
        close(v, gold[f"{name}.{k}"], 1e-6, 1e-7)
    if rc.get("plain"):
        out2 = rp.average_weights(w_loc, rc["idxs"], n_k, n_kg)
        for k, v in out2.items():
            close(v, gold[f"{name}.plain.{k}"], 1e-6, 1e-7)


def test_fedavg_survey_known_answer():
    """SURVEY.md §3.5 hand-checked values: 2 clients n=10/30, by_attr [[5,5,0],[5,10,15]], epoch 1 of 2."""
    G, r = 3, 4
    w = [{"lora_A": torch.full((2, r), 1.0), "lora_S": torch.full((G, r), 1.0)},
         {"lora_A": torch.full((2, r), 0.5), "lora_S": torch.tensor([[.5, .5, .5, .5]] * 3)}]
    w_g = {"lora_A": torch.full((2, r), 1.0), "lora_S": torch.full((G, r), 1.0)}
    out = rp.average_weights_ema(w_g, w, [0, 1], [10, 30], [[5, 5, 0], [5, 10, 15]], 1, 2, shared_half_s=True)
    bd = 0.999 * 0.5
    exp_A = (1 - bd) * (0.25 * 1.0 + 0.75 * 0.5) + bd * 1.0
    close(out["lora_A"], np.full((2, r), exp_A, dtype=np.float32), 1e-6, 1e-7)
    rows = np.array([0.5 * 1 + 0.5 * .5, (5 / 15) * 1 + (10 / 15) * .5, 0 * 1 + 1.0 * .5])
    exp_S = np.stack([np.full(G, rows.mean())] * 2 + [rows, rows], axis=1)
    close(out["lora_S"], ((1 - bd) * exp_S + bd).astype(np.float32), 1e-6, 1e-7)


# ----------------------------------------------------------------------------------------------- metrics
@pytest.mark.parametrize("name", list(recipes.METRIC_CASES))
def test_metrics_match_reference(name):
    rc = recipes.METRIC_CASES[name]
    gold = load("metrics.npz")
    prob, y, attrs = recipes.metric_inputs(rc)
    assert rp.compute_auc(prob, y) == pytest.approx(float(gold[f"{name}.auc"]), abs=1e-12)
    assert rp.compute_auc(prob[:, 1], y) == pytest.approx(float(gold[f"{name}.auc_binary"]), abs=1e-12)
    for a in range(attrs.shape[0]):
        assert rp.equity_scaled_accuracy(prob, y, attrs[a]) == pytest.approx(float(gold[f"{name}.esacc{a}"]), abs=1e-12)
        assert rp.equity_scaled_auc(prob, y, attrs[a]) == pytest.approx(float(gold[f"{name}.esauc{a}"]), abs=1e-12)
        groups = [e for e in np.unique(attrs[a]).astype(int) if e != -1]
        got = np.array([rp.compute_auc(prob[attrs[a] == e], y[attrs[a] == e]) for e in groups])
        np.testing.assert_allclose(got, gold[f"{name}.gauc{a}"], rtol=0, atol=1e-12)
    res = rp.comprehensive_scores(prob, y, attrs)
    assert res[0] == pytest.approx(float(gold[f"{name}.overall_acc"]), abs=1e-12)
    np.testing.assert_allclose(res[1], gold[f"{name}.esaccs"], atol=1e-12)
    assert res[2] == pytest.approx(float(gold[f"{name}.overall_auc"]), abs=1e-12)
    np.testing.assert_allclose(res[3], gold[f"{name}.esaucs"], atol=1e-12)
    np.testing.assert_allclose(res[8], gold[f"{name}.disparity"], atol=1e-12)


def test_mann_whitney_equals_sklearn():
    sk = pytest.importorskip("sklearn.metrics")
    rng = np.random.default_rng(0)
    for n in (50, 1000):
        y = rng.integers(0, 2, n)
        s = np.round(rng.random(n), 2).astype(np.float32)     # heavy ties
        gt, eq, P, Nn = rp.mann_whitney_counts(s, y == 1)
        assert rp.auc_from_counts(gt, eq, P, Nn) == pytest.approx(sk.roc_auc_score(y, s), abs=1e-12)
        # brute-force integer check
        pos, neg = s[y == 1], s[y == 0]
        assert gt == int((pos[:, None] > neg[None, :]).sum())
        assert eq == int((pos[:, None] == neg[None, :]).sum())


# ----------------------------------------------------------------------------------------------- whole model
@pytest.mark.parametrize("name", list(recipes.MODEL_CASES))
def test_custom_clip_forward_backward_matches_reference(name):
    rc = recipes.MODEL_CASES[name]
    gold = load("model.npz")
    keys = [str(k) for k in gold[f"{name}.keys"]]
    shapes = {k: tuple(int(v) for v in str(s).split(",") if v != "") for k, s in zip(keys, gold[f"{name}.shapes"])}
    params = recipes.model_params(rc, shapes)
    trainable = [k for k in params if rc["grad_filter"](k) and params[k].is_floating_point() and
                 "running_" not in k]
    for k in trainable:
        params[k].requires_grad_(True)
    image, label, attr = recipes.model_batch(rc)
    eot = torch.from_numpy(gold[f"{name}.eot"])
    logits = rp.custom_clip_forward(
        image, attr, params, eot, n_prompts=2, n_cls=2, ot=rc["ot"], vision_layers=rc["v_layers"],
        vision_heads=(rc["v_width"] * 32 // 64) if isinstance(rc["v_layers"], tuple) else rc["v_width"] // 64,
        text_layers=rc["t_layers"], text_heads=rc["t_heads"],
        scaling=rc["alpha"] / rc["rank"], lora_type=rc["lora_type"],
        dim_per_3d_slice=rc.get("dim_per_3d_slice") if rc["modality"] == "oct_bscans" else None)
    close(logits, gold[f"{name}.logits"], 2e-4, 2e-5)
    loss = torch.nn.functional.cross_entropy(logits, label)
    close(loss, gold[f"{name}.loss"], 1e-4, 1e-6)
    loss.backward()
    n_checked = 0
    for k in trainable:
        gk = f"{name}.grad.{k}"
        if gk in gold.files:
            g = params[k].grad
            ref = gold[gk]
            scale = max(1e-8, float(np.abs(ref).max()))
            np.testing.assert_allclose(g.numpy().reshape(ref.shape), ref, rtol=2e-3, atol=2e-4 * scale)
            n_checked += 1
    assert n_checked >= 7


@pytest.mark.parametrize("causal", [False, True])
def test_attention_core_is_pinned_to_torch_multihead_attention(causal):
    """The reference's blocks call torch's nn.MultiheadAttention (clip/model.py:350-352); the oracle's restatement of
    its core must reproduce the real module (same packed in_proj layout, scaling, mask handling)."""
    torch.manual_seed(3)
    L, B, C, H = 11, 3, 64, 4
    mha = torch.nn.MultiheadAttention(C, H)
    x = torch.randn(L, B, C)
    mask = torch.full((L, L), float("-inf")).triu_(1) if causal else None
    want = mha(x, x, x, need_weights=False, attn_mask=mask)[0]
    p = {"in_proj_weight": mha.in_proj_weight, "in_proj_bias": mha.in_proj_bias,
         "out_proj.weight": mha.out_proj.weight, "out_proj.bias": mha.out_proj.bias}
    got = rp._mha(x, p, "", H, mask)
    torch.testing.assert_close(got, want, rtol=1e-5, atol=1e-6)
