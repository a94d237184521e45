"""Host-side logic that needs no GPU: registry contract, module surgery / state-dict keys, flat specs, client
sampling, and the N>1 aggregation flow over gloo (world size 2) with the two CUDA kernel hooks replaced by the
oracle's arithmetic in a TEST-SIDE subclass."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

from fairfedmed_b200 import fed_utils, modules, registry
from fairfedmed_b200.clip_model import CustomCLIP
from fairfedmed_b200.federated import select_clients
from oracle import ref_port as rp
from tests.golden import recipes


def test_registry_contract():
    reg = registry.Registry("T")

    @reg.register()
    class A:
        pass

    assert reg.registered_names() == ["A"]
    with pytest.raises(KeyError):
        reg.register(A)
    reg.register(A, force=True)
    with pytest.raises(KeyError):
        reg.get("missing")
    import fairfedmed_b200.trainer  # noqa: F401
    assert "GLP_OT_SVLoRA" in registry.TRAINER_REGISTRY.registered_names()


def test_surgery_matches_reference_selection_and_keys():
    m = CustomCLIP(vision_layers=2, vision_width=128, text_layers=1, text_width=64, text_heads=2, embed_dim=64,
                   image_resolution=32)
    for n, p in m.named_parameters():
        p.requires_grad_("prompt_learner" in n)
    modules.apply_lora_to_model(m, True, rank=12, alpha=2, lora_type="FairLoRA", num_attrs=3)
    adapters = [n for n, mod in m.named_modules() if isinstance(mod, modules.FairLoRALinear)]
    assert adapters == [f"image_encoder.transformer.resblocks.{i}.mlp.{l}" for i in range(2) for l in ("c_fc", "c_proj")]
    sd = m.state_dict()
    k = "image_encoder.transformer.resblocks.0.mlp.c_fc."
    assert sd[k + "lora_A.weight"].shape == (128, 12)
    assert sd[k + "lora_S.weight"].shape == (3, 12)
    assert sd[k + "lora_B.weight"].shape == (12, 512)
    assert sd[k + "original_linear.weight"].shape == (512, 128)
    trainable = [n for n, p in m.named_parameters() if p.requires_grad]
    assert len(trainable) == 1 + 4 * 3          # ctx + 4 adapters x (A, S, B)
    # text tower and attention stay un-adapted
    assert not any("text_encoder" in n and "lora" in n for n in sd)
    assert not any(".attn." in n and "lora" in n for n in sd)


def test_fairlora_init_matches_oracle():
    layer = modules.FairLoRALinear(nn.Linear(64, 32), rank=12, alpha=2.0, num_attrs=3)
    torch.testing.assert_close(layer.lora_S.weight.data, rp.fairlora_init_S(3, 12))
    assert float(layer.lora_A.weight.abs().max()) == 0.0
    assert layer.scaling == pytest.approx(1 / 6)
    g = modules.FairLoRALinear(nn.Linear(64, 32), rank=8, alpha=2.0, global_s=True, num_attrs=2)
    assert g.lora_S_global.weight.shape == (8,)          # 1-D like upstream after reset_parameters


def test_vit_b16_trainable_parameter_count():
    """ViT-B/16, r=12, G=3 -> 4096 + 24 * 46116 = 1 110 880 trainable parameters."""
    r, G = 12, 3
    per_pair = (768 * r + G * r + r * 3072) + (3072 * r + G * r + r * 768)
    assert 2 * 4 * 512 + 12 * per_pair == 1_110_880


def test_flat_spec_kinds():
    rc = recipes.FEDAVG_CASES["frac"]
    w_g, _, _, _ = recipes.fedavg_inputs(rc)
    spec = fed_utils.build_spec(w_g, num_groups=rc["groups"])
    kinds = dict(zip(spec.keys, spec.kinds))
    assert kinds["image_encoder.transformer.resblocks.0.mlp.c_fc.lora_S.weight"] == 1
    assert kinds["image_encoder.transformer.resblocks.0.mlp.c_fc.lora_S_global.weight"] == 0
    assert kinds["prompt_learner.ctx"] == 0
    assert spec.numel == sum(v.numel() for v in w_g.values())
    flat = torch.cat([w_g[k].reshape(-1) for k in spec.keys])
    back = fed_utils.unpack(spec, flat)
    for k in spec.keys:
        torch.testing.assert_close(back[k], w_g[k])


def test_client_sampling_follows_reference_rule():
    rng = np.random.RandomState(1)
    assert select_clients(0, 3, 0.8, rng) == [0, 1, 2]
    for e in range(1, 6):
        sel = select_clients(e, 3, 0.8, rng)
        assert len(sel) == 2 and len(set(sel)) == 2


# ------------------------------------------------------------------------------------------- gloo, world size 2
class _OracleKernelAggregator(fed_utils.FederatedAggregator):
    """Test-side stand-in for the two CUDA kernels (same contract, oracle arithmetic on CPU tensors)."""

    def _scale(self, local_flat, w_scalar, w_group):
        out = local_flat.clone()
        for off, shp, kind in zip(self.spec.offsets, self.spec.shapes, self.spec.kinds):
            n = torch.Size(shp).numel()
            seg = out[off:off + n]
            if kind == 1:
                seg.view(shp).mul_(w_group[:, None])
            else:
                seg.mul_(w_scalar)
        return out

    def _epilogue(self, summed, prev, beta_decay, shared_half_s):
        out = summed.clone()
        for off, shp, kind in zip(self.spec.offsets, self.spec.shapes, self.spec.kinds):
            if kind == 1 and shared_half_s:
                t = out[off:off + torch.Size(shp).numel()].view(shp)
                r = shp[1]
                t[:, : r // 2] = t[:, : r // 2].mean(dim=0, keepdim=True)
        return (1 - beta_decay) * out + beta_decay * prev


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, case, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rc = recipes.FEDAVG_CASES[case]
    w_g, w_loc, n_k, n_kg = recipes.fedavg_inputs(rc)
    spec = fed_utils.build_spec(w_g, num_groups=rc["groups"])
    agg = _OracleKernelAggregator(spec)
    local = torch.cat([w_loc[rank][k].reshape(-1) for k in spec.keys])
    prev = torch.cat([w_g[k].reshape(-1) for k in spec.keys])
    out = agg.aggregate(local, prev, n_k[rank], n_kg[rank], rank in rc["idxs"], rc["epoch"], rc["max_epoch"],
                        shared_half_s=rc["shared_half_s"])
    q.put((rank, out.numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("case", ["two_clients"])
def test_sharded_aggregation_over_gloo_matches_oracle(case):
    rc = recipes.FEDAVG_CASES[case]
    world = rc["n_clients"]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, case, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    w_g, w_loc, n_k, n_kg = recipes.fedavg_inputs(rc)
    ref = rp.average_weights_ema(w_g, w_loc, rc["idxs"], n_k, n_kg, rc["epoch"], rc["max_epoch"],
                                 shared_half_s=rc["shared_half_s"])
    spec = fed_utils.build_spec(w_g, num_groups=rc["groups"])
    ref_flat = torch.cat([ref[k].reshape(-1) for k in spec.keys]).numpy()
    for r in range(world):
        np.testing.assert_allclose(outs[r], ref_flat, rtol=1e-6, atol=1e-7)
    np.testing.assert_array_equal(outs[0], outs[1])


def test_ops_host_helpers_without_gpu():
    """Pure host logic of the op layer: rank padding, attention support predicate, no-op join without pending writes,
    CPU tensors are rejected (no fallback)."""
    import pytest
    import torch
    from fairfedmed_b200 import _cabi, ops
    assert [ops.padded_rank(r) for r in (1, 12, 16, 17, 32)] == [16, 16, 16, 32, 32]
    with pytest.raises(_cabi.FfmError):
        ops.padded_rank(33)
    assert ops.attention_supported(768, 12, 197) and ops.attention_supported(512, 8, 77)
    assert not ops.attention_supported(768, 8, 197) and not ops.attention_supported(768, 12, 209)
    ops.join_direct_grad_writes()                       # nothing pending: must not touch CUDA
    assert ops._direct_grad(torch.zeros(2)) is None
    t = torch.zeros(2)
    t._ffm_direct_grad = torch.ones(2)
    assert ops._direct_grad(t) is t._ffm_direct_grad
    with pytest.raises(_cabi.FfmError, match="no CPU fallback"):
        ops.patchify_normalize(torch.zeros(1, 3, 16, 16), torch.zeros(3), torch.ones(3), 16, True)
    with pytest.raises(_cabi.FfmError, match="no CPU fallback"):
        ops.attention(torch.zeros(1, 4, 192, dtype=torch.bfloat16), 1, False, True)


def test_adapt_attention_opt_in_adds_adapters_without_touching_reference_keys():
    """North-star wording "every attention and MLP linear": opt-in FairLoRA on in_proj / out_proj of the image tower.
    Off by default (the reference adapts the MLP only): then the state-dict keys are the reference's; on, only
    adapter tensors are added and the frozen attention parameters keep their single key."""
    def build(flag):
        m = CustomCLIP(vision_layers=2, vision_width=128, text_layers=1, text_width=64, text_heads=2, embed_dim=64,
                       image_resolution=32)
        for n, p in m.named_parameters():
            p.requires_grad_("prompt_learner" in n)
        modules.apply_lora_to_model(m, True, rank=12, alpha=2, lora_type="FairLoRA", num_attrs=3, adapt_attention=flag)
        return m
    off, on = build(False), build(True)
    k_off, k_on = set(off.state_dict()), set(on.state_dict())
    assert not any("attn_in_lora" in k or "attn_out_lora" in k for k in k_off)
    extra = k_on - k_off
    assert extra == {f"image_encoder.transformer.resblocks.{i}.{a}.{w}.weight" for i in range(2)
                     for a in ("attn_in_lora", "attn_out_lora") for w in ("lora_A", "lora_S", "lora_B")}
    sd = on.state_dict()
    assert sd["image_encoder.transformer.resblocks.0.attn_in_lora.lora_A.weight"].shape == (128, 12)
    assert sd["image_encoder.transformer.resblocks.0.attn_in_lora.lora_B.weight"].shape == (12, 384)
    assert sd["image_encoder.transformer.resblocks.1.attn_out_lora.lora_B.weight"].shape == (12, 128)
    trainable = [n for n, p in on.named_parameters() if p.requires_grad]
    assert len(trainable) == 1 + 2 * 4 * 3                       # ctx + 2 blocks x (c_fc, c_proj, in, out) x (A, S, B)
    assert not any(p.requires_grad for n, p in on.named_parameters() if ".attn.in_proj" in n or ".attn.out_proj" in n)
    with pytest.raises(NotImplementedError):
        modules.apply_lora_to_model(build(False), True, rank=4, lora_type="LoRA", adapt_attention=True)


def test_resnet_glue_falls_back_to_the_library_off_gpu():
    """The ResNet trunk's glue ops (BatchNorm + ReLU, average pooling, relu(out + identity)) use the package's kernels only for
    CUDA tensors in the layouts they were built for; anything else (CPU tensors here) must take the library path and give
    torch's own results — construction / state-dict handling of the model happens on the CPU."""
    import torch
    import torch.nn as nn
    import torch.nn.functional as F
    from fairfedmed_b200 import ops, resnet_model as rm
    torch.manual_seed(0)
    x = torch.randn(2, 8, 6, 6)
    bn = nn.BatchNorm2d(8)
    ref_bn = nn.BatchNorm2d(8)
    ref_bn.load_state_dict(bn.state_dict())
    y = rm._bn(bn, x, True)
    torch.testing.assert_close(y, F.relu(ref_bn(x)))
    torch.testing.assert_close(bn.running_mean, ref_bn.running_mean)
    assert int(bn.num_batches_tracked) == 1
    torch.testing.assert_close(rm._AvgPool2d(2)(x), F.avg_pool2d(x, 2))
    torch.testing.assert_close(ops.add_relu(x, -0.5 * x), torch.relu(0.5 * x))
    assert not ops.batchnorm_relu_supported(x) and not ops.avgpool_nhwc_supported(x, 2)
