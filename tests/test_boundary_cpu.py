"""Drop-in boundary against the REAL reference control plane (CPU, needs /root/reference — skipped elsewhere).

The reference's own `federated_main.setup_cfg` builds the cfg from its own yaml files and launch-script arguments
(scripts/fairfedlora_fairfedmed.sh); the B200 trainer is registered in Dassl's real TRAINER_REGISTRY under the same
name with force=True (Dassl/dassl/utils/registry.py:36-43) and built through Dassl's `build_trainer(cfg)`
(Dassl/dassl/engine/build.py:14-20), which only passes cfg.  What is asserted is the surface federated_main.py uses
(:183-207, :616-690): ctor(cfg), fed_before_train, train / test signatures, fed_train_loader_x_dict[i].dataset (len,
count_by_attribute), dm.dataset.classnames, model.state_dict / load_state_dict with the reference's key names — and
that weights + prompt buffers loaded from a CLIP state dict equal what the reference's CustomCLIP holds.
The constructor is run with `_require_cuda = False` (tensors on the CPU, no kernel is launched at construction).
"""
from __future__ import annotations

import argparse
import inspect

import pytest
import torch

from oracle import shim

pytestmark = pytest.mark.skipif(not shim.available(), reason="reference tree not present")


def _args(**over):
    """The argparse defaults of federated_main.py:790-877 overridden by scripts/fairfedlora_fairfedmed.sh:7-41,58-90."""
    ref = shim.REF_ROOT
    a = dict(model="FedOTPLoRA", trainer="GLP_OT_SVLoRA", round=50, stepsize=200, num_users=3, frac=0.8, lr=0.001,
             gamma=0.1, train_batch_size=32, test_batch_size=100, seed=1, mu=0.5, disease_type="heart.attack", iid=False,
             num_shots=2, useall=True, partition="noniid-labeldir100", beta=0.3, imbalance_train=False,
             split_client=False, num_domain=4, attribute_type="language",
             attributes=["race", "language", "ethnicity"], modality_type="slo_fundus", dim_per_3d_slice=16,
             input_no_transform=False, n_ctx=4, num_prompt=2, avg_prompt=1, ctx_init=False, OT="None", top_percent=0.8,
             eps=0.1, thresh=1e-3, max_iter=100, unfreeze_image_encoder=True, unfreeze_text_encoder=False, lora_rank=12,
             lora_alpha=2.0, lora_type="FairLoRA", lora_local_s=False, shared_half_s=True, lora_global_s=False,
             lambda_fairness=0.0, idxs_users_train=[], idxs_users_test=[], disable_attr=False, logdir="./logs/",
             root="DATA/", output_dir="output/test", config_file=f"{ref}/configs/trainers/GLP_OT/vit_b16_oph.yaml",
             dataset_config_file=f"{ref}/configs/datasets/fairfedmed.yaml", resume=None, transforms=None, backbone="",
             head="", eval_only=False, model_dir="", load_epoch=None, no_train=False, opts=[])
    a.update(over)
    return argparse.Namespace(**a)


@pytest.fixture(scope="module")
def ref_env():
    shim.install()
    import federated_main as FM
    from Dassl.dassl.engine import TRAINER_REGISTRY, build_trainer
    from fairfedmed_b200.trainer import GLP_OT_SVLoRA as B200Trainer
    TRAINER_REGISTRY.register(B200Trainer, force=True)              # INTEGRATION.md option (c)
    return FM, TRAINER_REGISTRY, build_trainer, B200Trainer


class _FakeDataset:
    def __init__(self, n, counts):
        self._n, self._c = n, counts

    def __len__(self):
        return self._n

    def count_by_attribute(self, name):
        return list(self._c)


class _FakeDM:
    """Shape of Dassl's DataManager as the trainer and federated_main.py use it (data_manager.py:62-201)."""

    def __init__(self, cfg):
        self.dataset = type("D", (), {"classnames": ["NOT Glaucoma", "Glaucoma"],
                                      "lab2cname": {0: "NOT Glaucoma", 1: "Glaucoma"}, "num_classes": 2})()
        mk = lambda n, c: type("L", (), {"dataset": _FakeDataset(n, c), "__len__": lambda s: 1})()
        self.fed_train_loader_x_dict = {k: mk(10 + k, [3, 3, 4 + k]) for k in range(cfg.DATASET.USERS)}
        self.fed_test_loader_x_dict = {k: mk(5, [2, 2, 1]) for k in range(cfg.DATASET.USERS)}


def _tiny_clip(CM, vision_layers=2, width=128):
    torch.manual_seed(0)
    dd = {"trainer": "GLP_OT", "vision_depth": 0, "language_depth": 0, "vision_ctx": 0, "language_ctx": 0}
    heads = 2
    return CM.CLIP(64, 224, vision_layers, width, 16, 77, 49408, 128, heads, 2, dd).float()


def test_reference_cfg_builds_the_b200_trainer_through_dassl(ref_env, monkeypatch):
    FM, REG, build_trainer, B200 = ref_env
    cfg = FM.setup_cfg(_args())
    assert cfg.TRAINER.NAME == "GLP_OT_SVLoRA" and cfg.MODEL.BACKBONE.NAME == "ViT-B/16"
    assert not hasattr(cfg, "MODEL_ARCH")                       # the reference's cfg knows nothing about tower shapes
    assert REG.get("GLP_OT_SVLoRA") is B200

    # without data or weights the constructor must fail loudly, not train random weights on random data
    monkeypatch.setattr(B200, "_require_cuda", False, raising=False)
    monkeypatch.setattr(B200, "data_manager_factory", None)
    import Dassl.dassl.data as DD
    monkeypatch.setattr(DD, "DataManager", lambda cfg: (_ for _ in ()).throw(FileNotFoundError("no dataset")))
    with pytest.raises(FileNotFoundError):
        build_trainer(cfg)
    monkeypatch.setattr(B200, "data_manager_factory", staticmethod(lambda c: _FakeDM(c)))
    with pytest.raises(RuntimeError, match="pretrained CLIP weights"):
        build_trainer(cfg)

    # with the reference's pieces plugged in: CLIP weights (random-init tiny CLIP here), clip.tokenize
    T, CM, _, _ = shim.modules()
    import clip as ref_clip
    clip_model = _tiny_clip(CM)
    monkeypatch.setattr(B200, "clip_loader", staticmethod(lambda c: clip_model))
    monkeypatch.setattr(B200, "tokenize", staticmethod(ref_clip.clip.tokenize))
    tr = build_trainer(cfg)                                     # Dassl's own factory: Trainer(cfg)
    assert isinstance(tr, B200)

    # ---- the surface federated_main.py touches
    for name in ("fed_before_train", "fed_after_train", "train", "test", "model_inference", "parse_batch_train",
                 "parse_batch_test", "forward_backward", "run_epoch", "before_train", "after_train", "before_epoch",
                 "after_epoch", "update_lr", "get_current_lr", "set_model_mode", "save_model_with_grad"):
        assert callable(getattr(tr, name)), name
    sig = inspect.signature(tr.train)
    assert {"idx", "global_epoch", "is_fed", "is_last_client"} <= set(sig.parameters)
    assert {"idx", "current_epoch"} <= set(inspect.signature(tr.test).parameters)
    assert tr.dm.dataset.classnames == ["NOT Glaucoma", "Glaucoma"]
    assert len(tr.fed_train_loader_x_dict[1].dataset) == 11
    assert tr.fed_train_loader_x_dict[2].dataset.count_by_attribute("language") == [3, 3, 6]
    assert tr.max_epoch == 1 and tr.num_groups == 3 and tr.current_lr() == pytest.approx(1e-3)
    tr.fed_before_train()

    # ---- same model as the reference builds from the same cfg + CLIP weights: keys, shapes, frozen values, buffers
    ref_model = T.CustomCLIP(cfg, tr.classnames, clip_model)
    T.apply_lora_to_model(ref_model, True, rank=12, alpha=2.0, lora_type="FairLoRA", global_s=False, num_attrs=3)
    ref_sd, sd = ref_model.state_dict(), tr.model.state_dict()
    assert set(sd) == set(ref_sd)
    for k, v in ref_sd.items():
        assert tuple(sd[k].shape) == tuple(v.shape), k
        if "lora_" in k or k == "prompt_learner.ctx":
            continue                                            # randomly initialised on both sides
        torch.testing.assert_close(sd[k].float().cpu(), v.float(), rtol=0, atol=0, msg=k)
    assert torch.equal(tr.model.prompt_learner.eot_index.cpu(), ref_model.tokenized_prompts.argmax(-1))
    # `prompt_learner.ctx` is sliced by name, `lora_S` keys are filtered by substring (federated_main.py:624-625)
    assert sd["prompt_learner.ctx"].shape == (2, 4, 128)
    assert sum("lora_S" in k for k in sd) == 2 * 2
    # load_state_dict(strict=False) with the reference's dict must copy INTO the flat buffer (objects persist)
    before = {n: p for n, p in tr.model.named_parameters()}
    new = {k: torch.full_like(v, 0.25) for k, v in ref_sd.items() if "lora_A" in k}
    tr.model.load_state_dict(new, strict=False)
    assert all(before[n] is p for n, p in tr.model.named_parameters())
    k0 = next(iter(new))
    off = tr.flat_spec.offsets[tr.flat_spec.keys.index(k0)]
    assert float(tr.flat_all[off]) == 0.25


def test_backbone_name_selects_the_architecture(ref_env):
    from fairfedmed_b200 import clip_weights as cw
    assert cw.arch_for_backbone("ViT-B/16")["VISION_WIDTH"] == 768 and cw.arch_for_backbone("ViT-B/16")["PATCH"] == 16
    assert cw.arch_for_backbone("RN50")["VISION_LAYERS"] == (3, 4, 6, 3) and cw.arch_for_backbone("RN50")["EMBED"] == 1024
    _, CM, _, _ = shim.modules()
    sd = _tiny_clip(CM, vision_layers=3, width=192).state_dict()
    a = cw.arch_from_clip_state_dict(sd)
    assert (a["VISION_LAYERS"], a["VISION_WIDTH"], a["PATCH"], a["EMBED"], a["TEXT_WIDTH"], a["TEXT_LAYERS"]) == \
        (3, 192, 16, 64, 128, 2)
    with pytest.raises(KeyError):
        cw.arch_for_backbone("ViT-L/14@336px")
