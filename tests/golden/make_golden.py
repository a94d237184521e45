"""Generate the committed golden vectors by running the REAL reference (imported via oracle/shim.py).

Run once in the build container (needs /root/reference):   python tests/golden/make_golden.py
Outputs small .npz fixtures next to this file.  Inputs are produced by tests/golden/recipes.py from seeds so
the fixtures only need to hold the reference *outputs* (plus small inputs where convenient).
"""
from __future__ import annotations

import copy
import sys
from pathlib import Path

import numpy as np
import torch
import torch.nn as nn

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))

from oracle import shim  # noqa: E402
from tests.golden import recipes  # noqa: E402


def to_np(d):
    return {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in d.items()}


def golden_fairlora(T):
    out = {}
    for name, rc in recipes.FAIRLORA_CASES.items():
        t = recipes.fairlora_inputs(rc)
        lin = nn.Linear(rc["c_in"], rc["c_out"])
        with torch.no_grad():
            lin.weight.copy_(t["W"])
            lin.bias.copy_(t["bias"])
        if rc["kind"] == "FairLoRA":
            mod = T.FairLoRALinear(lin, rank=rc["rank"], alpha=rc["alpha"], global_s=rc["global_s"],
                                   num_attrs=rc["groups"])
        elif rc["kind"] == "SVLoRA":
            mod = T.SVLoRALinear(lin, rank=rc["rank"], alpha=rc["alpha"], global_s=rc["global_s"])
        else:
            mod = T.LoRALinear(lin, rank=rc["rank"], alpha=rc["alpha"])
        out[f"{name}.init_S"] = mod.lora_S.weight.detach().clone() if hasattr(mod, "lora_S") else torch.zeros(1)
        with torch.no_grad():
            mod.lora_A.weight.copy_(t["A"])
            mod.lora_B.weight.copy_(t["B"])
            if rc["kind"] == "FairLoRA":
                mod.lora_S.weight.copy_(t["S"])
                if rc["global_s"]:
                    mod.lora_S_global.weight.copy_(t["S_global"])
            elif rc["kind"] == "SVLoRA":
                mod.lora_S.weight.data = t["S"].reshape(-1).clone()      # reset_parameters leaves it 1-D upstream
                if rc["global_s"]:
                    mod.lora_S_global.weight.data = t["S_global"].reshape(-1).clone()
        x = t["x"].clone().requires_grad_(True)
        y = mod(x, t["attr"]) if rc["kind"] == "FairLoRA" else mod(x)
        (y * t["dy"]).sum().backward()
        out[f"{name}.y"] = y
        out[f"{name}.dx"] = x.grad
        out[f"{name}.dA"] = mod.lora_A.weight.grad
        out[f"{name}.dB"] = mod.lora_B.weight.grad
        if hasattr(mod, "lora_S"):
            out[f"{name}.dS"] = mod.lora_S.weight.grad
        if rc.get("global_s") and rc["kind"] == "FairLoRA":
            out[f"{name}.dS_global"] = mod.lora_S_global.weight.grad
        if rc["kind"] == "FairLoRA" and rc.get("merged"):
            out[f"{name}.merged_w"] = mod.weight(t["x"], t["attr"])
    np.savez_compressed(HERE / "fairlora.npz", **to_np(out))
    print("fairlora.npz", len(out), "arrays")


def golden_sinkhorn(T):
    out = {}
    cfg = shim.make_cfg(ot="Sinkhorn")
    for name, rc in recipes.SINKHORN_CASES.items():
        K, u, v = recipes.sinkhorn_inputs(rc)
        holder = T.CustomCLIP.__new__(T.CustomCLIP)          # only thresh / max_iter are read by the two methods
        holder.thresh = rc["thresh"]
        holder.max_iter = rc["max_iter"]
        if rc["mode"] == "Sinkhorn":
            plan = T.CustomCLIP.Sinkhorn(holder, K, u, v)
        else:
            plan = T.CustomCLIP.entropic_COT_fast(holder, u, v, K, 0.01, numItermax=rc["max_iter"])
        out[f"{name}.T"] = plan
    np.savez_compressed(HERE / "sinkhorn.npz", **to_np(out))
    print("sinkhorn.npz", len(out), "arrays")
    del cfg


def golden_fedavg(FU):
    out = {}
    for name, rc in recipes.FEDAVG_CASES.items():
        w_g, w_loc, n_k, n_kg = recipes.fedavg_inputs(rc)
        res = FU.average_weights_EMA(copy.deepcopy(w_g), copy.deepcopy(w_loc), rc["idxs"], n_k, n_kg, rc["epoch"],
                                     rc["max_epoch"], shared_half_s=rc["shared_half_s"])
        for k, v in res.items():
            out[f"{name}.{k}"] = v
        if rc.get("plain"):
            res2 = FU.average_weights(copy.deepcopy(w_loc), rc["idxs"], n_k, n_kg)
            for k, v in res2.items():
                out[f"{name}.plain.{k}"] = v
    np.savez_compressed(HERE / "fedavg.npz", **to_np(out))
    print("fedavg.npz", len(out), "arrays")


def golden_metrics(EM):
    out = {}
    for name, rc in recipes.METRIC_CASES.items():
        prob, y, attrs = recipes.metric_inputs(rc)
        out[f"{name}.auc"] = EM.compute_auc(prob, y)
        out[f"{name}.auc_binary"] = EM.compute_auc(prob[:, 1], y)
        for a in range(attrs.shape[0]):
            out[f"{name}.esacc{a}"] = EM.equity_scaled_accuracy(prob, y, attrs[a])
            out[f"{name}.esauc{a}"] = EM.equity_scaled_AUC(prob, y, attrs[a])
            groups = [e for e in np.unique(attrs[a]).astype(int) if e != -1]
            out[f"{name}.gauc{a}"] = np.array([EM.compute_auc(prob[attrs[a] == e], y[attrs[a] == e]) for e in groups])
        res = EM.evalute_comprehensive_perf_scores(prob, y, attrs)
        out[f"{name}.overall_acc"] = res[0]
        out[f"{name}.esaccs"] = res[1]
        out[f"{name}.overall_auc"] = res[2]
        out[f"{name}.esaucs"] = res[3]
        out[f"{name}.disparity"] = res[8]
    np.savez_compressed(HERE / "metrics.npz", **to_np(out))
    print("metrics.npz", len(out), "arrays")


def golden_model(T, CM):
    """Whole CustomCLIP forward/backward on a shrunken CLIP (same code path, tiny widths)."""
    out = {}
    for name, rc in recipes.MODEL_CASES.items():
        cfg = shim.make_cfg(modality=rc["modality"], ot=rc["ot"], dim_per_3d_slice=rc.get("dim_per_3d_slice", 8))
        cfg.INPUT.SIZE = (rc["res"], rc["res"])
        torch.manual_seed(0)
        dd = {"trainer": "GLP_OT", "vision_depth": 0, "language_depth": 0, "vision_ctx": 0, "language_ctx": 0}
        clip_model = CM.CLIP(rc["embed"], rc["res"], rc["v_layers"], rc["v_width"], 16, 77, 49408, rc["t_width"],
                             rc["t_heads"], rc["t_layers"], dd).float()
        model = T.CustomCLIP(cfg, list(recipes.CLASSNAMES), clip_model)
        for n_, p_ in model.named_parameters():
            p_.requires_grad_("prompt_learner" in n_ or "proj_per_3d_slice" in n_ or
                              (rc.get("train_bn", False) and (".bn" in n_ or "downsample.1" in n_)))
        T.apply_lora_to_model(model, True, rank=rc["rank"], alpha=rc["alpha"], lora_type=rc["lora_type"],
                              global_s=False, num_attrs=rc["groups"])
        eot = model.tokenized_prompts.argmax(dim=-1)
        sd_ref = model.state_dict()
        params = recipes.model_params(rc, {k: tuple(v.shape) for k, v in sd_ref.items()})
        missing = model.load_state_dict(params, strict=True)
        image, label, attr = recipes.model_batch(rc)
        logits = model(image, attr)
        loss = torch.nn.functional.cross_entropy(logits, label)
        loss.backward()
        out[f"{name}.eot"] = eot
        out[f"{name}.logits"] = logits
        out[f"{name}.loss"] = loss
        out[f"{name}.keys"] = np.array(list(sd_ref.keys()))
        out[f"{name}.shapes"] = np.array([",".join(map(str, v.shape)) for v in sd_ref.values()])
        for n_, p_ in model.named_parameters():
            if p_.grad is not None and rc["grad_filter"](n_):
                out[f"{name}.grad.{n_}"] = p_.grad
        del missing
    np.savez_compressed(HERE / "model.npz", **to_np(out))
    print("model.npz", len(out), "arrays")


def main():
    T, CM, FU, EM = shim.modules()
    torch.set_num_threads(8)
    golden_fairlora(T)
    golden_sinkhorn(T)
    golden_fedavg(FU)
    golden_metrics(EM)
    golden_model(T, CM)


if __name__ == "__main__":
    main()
