"""GPU parity of the ViT input-side kernels (scope row f3) and of the training-step plumbing added around them:
patchify+normalise (bit-exact vs the reference's expression order), class/positional embedding + ln_pre + ln_1
(vs an fp32 torch restatement of clip/model.py:434-440,:354), direct adapter-gradient writes (same result as the
autograd-accumulated path) and the side-stream schedule (text tower, hoisted adapter preparation: same result as the
single-stream schedule, eagerly and replayed from the captured step graph)."""
from __future__ import annotations

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("bp,res,patch", [(3, 32, 16), (2, 224, 16), (5, 64, 8)])
def test_patchify_normalize_is_bit_exact(bp, res, patch):
    from fairfedmed_b200 import ops
    from fairfedmed_b200.clip_model import PIXEL_MEAN, PIXEL_STD
    g = torch.Generator().manual_seed(bp * 1000 + res)
    img = torch.randint(0, 256, (bp, 3, res, res), generator=g).float().to(DEV)
    mean = torch.tensor(PIXEL_MEAN, device=DEV)
    std = torch.tensor(PIXEL_STD, device=DEV)
    out = ops.patchify_normalize(img, mean, std, patch, True)
    # trainers/GLP_OT_SVLoRA.py:679-693 then clip/model.py:431-433 (stride = kernel: im2col of the cast image)
    x = ((img / 255.0) - mean.view(1, 3, 1, 1)) / std.view(1, 3, 1, 1)
    gh = res // patch
    ref = x.to(torch.bfloat16).view(bp, 3, gh, patch, gh, patch).permute(0, 2, 4, 1, 3, 5).reshape(bp, gh * gh, -1)
    assert out.shape == ref.shape and out.dtype == torch.bfloat16
    assert torch.equal(out, ref)
    # div255 = 0: the caller already scaled the image (OCT min-max path)
    out2 = ops.patchify_normalize(img / 255.0, mean, std, patch, False)
    assert torch.equal(out2, ref)


@pytest.mark.parametrize("bp,G,C", [(4, 196, 768), (3, 16, 256), (2, 49, 512), (1, 4, 1024)])
def test_vit_embed_ln_matches_torch(bp, G, C):
    from fairfedmed_b200 import ops
    g = torch.Generator().manual_seed(G * 7 + C)
    pe = torch.randn(bp, G, C, generator=g).to(DEV).to(torch.bfloat16)
    cls = (C ** -0.5 * torch.randn(C, generator=g)).to(DEV)
    pos = (C ** -0.5 * torch.randn(G + 1, C, generator=g)).to(DEV)
    gp, bpre = (1 + 0.1 * torch.randn(C, generator=g)).to(DEV), (0.1 * torch.randn(C, generator=g)).to(DEV)
    g1, b1 = (1 + 0.1 * torch.randn(C, generator=g)).to(DEV), (0.1 * torch.randn(C, generator=g)).to(DEV)
    x0, h0 = ops.vit_embed_ln(pe, cls, pos, gp, bpre, g1, b1, 1e-5, 1e-5)
    tok = torch.cat([cls.expand(bp, 1, C), pe.float()], dim=1) + pos
    x_ref = F.layer_norm(tok, (C,), gp, bpre, 1e-5)
    # bf16 output rounding: 2^-8 relative + a little absolute slack for values near zero
    assert torch.allclose(x0.float(), x_ref, rtol=2 ** -7, atol=2e-3)
    # ln_1 normalises the STORED bf16 stream: compare against LayerNorm of x0 itself (tight) and of the fp32 chain (loose)
    h_ref = F.layer_norm(x0.float(), (C,), g1, b1, 1e-5)
    assert torch.allclose(h0.float(), h_ref, rtol=2 ** -7, atol=2e-3)
    assert torch.allclose(h0.float(), F.layer_norm(x_ref, (C,), g1, b1, 1e-5), rtol=0, atol=6e-2)


def _small_trainer(direct: bool, overlap: bool):
    import bench
    import fairfedmed_b200.trainer  # noqa: F401  (registers GLP_OT_SVLoRA)
    from fairfedmed_b200.registry import build_trainer
    cfg = bench.make_cfg(1, 4, "Sinkhorn", bench.CONFIGS[2])
    cfg.MODEL_ARCH.VISION_LAYERS = 2
    cfg.MODEL_ARCH.TEXT_LAYERS = 2
    cfg.INPUT.SIZE = (64, 64)
    tr = build_trainer(cfg)
    tr.sync_metrics = False
    tr.step_auc = False
    tr.model.check_nan = False
    tr.model.overlap_text = overlap                                    # text tower on a side stream
    tr.model.image_encoder.transformer.hoist_adapter_prep = overlap    # s_eff / adapter tiles of all blocks up front
    tr.batch_idx, tr.num_batches = 0, 10 ** 9
    if not direct:
        for p in tr.model.parameters():
            if hasattr(p, "_ffm_direct_grad"):
                del p._ffm_direct_grad
    g = torch.Generator().manual_seed(3)
    with torch.no_grad():
        for n_, p_ in tr.model.named_parameters():
            if "lora_A" in n_:
                p_.copy_((0.02 * torch.randn(p_.shape, generator=g)).to(DEV))
    batch = {"img": torch.randint(0, 256, (4, 1, 64, 64), generator=g).float().repeat(1, 3, 1, 1),
             "label": torch.tensor([0, 1, 0, 1]), "attrs": torch.randint(0, 3, (4, 1), generator=g)}
    return tr, batch


def test_direct_gradients_and_side_stream_do_not_change_the_step():
    """Same seed, same batch, two steps: gradients and parameters after the steps must agree whether the adapter
    gradients are written in place or accumulated by autograd, and whether the text tower runs on a side stream."""
    from fairfedmed_b200 import ops
    results = []
    for direct, overlap in ((False, False), (True, False), (True, True)):
        tr, batch = _small_trainer(direct, overlap)
        ops.PARAMS_ON_SIDE_STREAM = overlap          # adapter gradients behind the dX GEMMs on their own stream
        try:
            for _ in range(2):
                tr.forward_backward(batch)
        finally:
            ops.PARAMS_ON_SIDE_STREAM = True
        torch.cuda.synchronize()
        assert bool(torch.isfinite(tr.flat_params).all())
        results.append((tr.flat_params.clone(), tr.flat_grads.clone()))
        names = [n for n, p in tr.model.named_parameters() if hasattr(p, "_ffm_direct_grad")]
        assert (len(names) > 0) == direct
    base_p, base_g = results[0]
    gmax = float(base_g.abs().max())
    assert gmax > 0
    # not bit-equal by construction: the library attention backward accumulates dQ with atomics, so two runs differ in
    # the last bits; a lost / doubled / misplaced gradient would be off by O(1) of the largest entry
    for p, g in results[1:]:
        assert float((g - base_g).abs().max()) <= 2e-3 * gmax
        assert float((p - base_p).abs().max()) <= 1e-5
    # the singular-value gradients are orders of magnitude smaller than the rest: check them on their own scale
    tr, _ = _small_trainer(True, True)
    off = 0
    checked = 0
    for n, prm in tr.model.named_parameters():
        if not prm.requires_grad:
            continue
        k = prm.numel()
        if "lora_S" in n:
            ref_s = base_g[off:off + k]
            smax = float(ref_s.abs().max())
            assert smax > 0, n
            for _, g in results[1:]:
                assert float((g[off:off + k] - ref_s).abs().max()) <= 2e-2 * smax, n
            checked += 1
        off += k
    assert checked == 4


def test_graphed_step_with_side_stream_matches_eager():
    """The captured step (fork/join of the text stream inside the graph) reproduces the eager losses; the capture's
    warm-up steps are rolled back (parameters, momentum), so step k of both trainers sees the same state."""
    tr_e, batch = _small_trainer(True, True)
    tr_g, _ = _small_trainer(True, True)
    eager = [tr_e.forward_backward(batch)["loss"].item() for _ in range(3)]
    first = tr_g.forward_backward(batch)["loss"].item()          # first step eagerly
    dev_batch = {k: v.to(DEV) for k, v in batch.items()}
    before = tr_g.get_flat().clone()
    tr_g.capture_step_graph(dev_batch, warmup=2)
    assert torch.equal(tr_g.get_flat(), before), "capture must not move the parameters"
    graphed = [tr_g.forward_backward_graphed(dev_batch)["loss"].item() for _ in range(2)]
    assert first == pytest.approx(eager[0], abs=1e-5)
    assert graphed[0] == pytest.approx(eager[1], abs=1e-4)
    assert graphed[1] == pytest.approx(eager[2], abs=1e-4)
    assert float((tr_g.flat_params - tr_e.flat_params).abs().max()) <= 1e-5


def test_train_replays_the_graph_and_matches_the_eager_epoch():
    """The PUBLIC train(idx=...) path: `step_metrics = "epoch"` (captured graph, staged batches, deferred read-back) and
    `"step"` (eager, python floats per step like the reference) give the same per-step losses / accuracies / training
    AUCs and the same parameters over two client-epochs — including a StepLR change between them, which the graph
    picks up from device memory without re-capture, and a graph captured from the very first step (zero momentum)."""
    import fairfedmed_b200.trainer  # noqa: F401
    from fairfedmed_b200.data import CachedLoader
    trs = []
    for mode in ("step", "epoch"):
        tr, _ = _small_trainer(True, True)
        tr.cfg.OPTIM.STEPSIZE = (2,)                 # StepLR(step_size=2): the rate drops after the first client-epoch
        tr.cfg.OPTIM.GAMMA = 0.5
        tr.step_metrics, tr.sync_metrics, tr.step_auc = mode, True, True
        batches = list(tr.fed_train_loader_x_dict[0])
        assert len(batches) >= 2
        tr.fed_train_loader_x_dict[0] = CachedLoader(batches, len(batches), tr.fed_train_loader_x_dict[0].dataset)
        logs = []
        for ep in range(2):
            last = tr.train(idx=0, global_epoch=ep, is_fed=True, is_last_client=True)
            logs.append(tr.last_epoch_summaries if mode == "epoch" else [last])
        trs.append((tr, logs))
    (te, le), (tg, lg) = trs
    assert tg._graph is not None and te._graph is None
    assert te.sched_steps == tg.sched_steps == 4 and te.current_lr() == pytest.approx(0.25 * te.base_lr)
    for ep in range(2):
        assert lg[ep][-1]["loss"] == pytest.approx(le[ep][-1]["loss"], abs=2e-4)
        assert lg[ep][-1]["acc"] == pytest.approx(le[ep][-1]["acc"], abs=1e-4)
        assert lg[ep][-1]["auc"] == pytest.approx(le[ep][-1]["auc"], abs=1e-6)
    assert float((tg.flat_params - te.flat_params).abs().max()) <= 2e-5


@pytest.mark.gpu
@pytest.mark.parametrize("shape,patch", [((6, 3, 32, 32), 8), ((5, 3, 224, 224), 16)])
def test_oct_minmax_patchify_forward_and_backward_match_autograd(shape, patch):
    """OCT input side after the slice projection (trainers/GLP_OT_SVLoRA.py:686-693): per-slice min-max scaling, mean / std,
    cast, im2col — fused forward and backward against autograd over the reference's own expressions (fp32)."""
    from fairfedmed_b200 import ops
    torch.manual_seed(5)
    bp, c, h, w = shape
    y = torch.randn(shape, device="cuda:0") * 3.0
    y[0, 0, 0, :4] = y[0].max() + 1.0           # ties at the maximum: amax splits the gradient evenly
    y[1, 2, 3, 5:7] = y[1].min() - 1.0          # ties at the minimum
    mean = torch.tensor([0.48, 0.46, 0.41], device="cuda:0")
    std = torch.tensor([0.27, 0.26, 0.28], device="cuda:0")
    w_out = torch.randn(bp, (h // patch) * (w // patch), c * patch * patch, device="cuda:0")

    y1 = y.clone().requires_grad_(True)
    got = ops.oct_minmax_patchify(y1, mean, std, patch)
    (got.float() * w_out).sum().backward()

    y2 = y.clone().requires_grad_(True)
    lo = y2.amin(dim=(1, 2, 3), keepdim=True)
    hi = y2.amax(dim=(1, 2, 3), keepdim=True)
    z = (y2 - lo) / (hi - lo + 1e-5)
    z = (z - mean.view(1, 3, 1, 1)) / std.view(1, 3, 1, 1)
    ref = z.reshape(bp, c, h // patch, patch, w // patch, patch).permute(0, 2, 4, 1, 3, 5).reshape(got.shape)
    # the fused backward sees the bf16-rounded upstream gradient, so hand the reference the same one
    (ref * w_out.to(torch.bfloat16).float()).sum().backward()

    assert torch.equal(got, ref.to(torch.bfloat16))
    scale = float(y2.grad.abs().max())
    assert float((y1.grad - y2.grad).abs().max()) <= 2e-4 * scale
    # deterministic: same bits on a second run
    y3 = y.clone().requires_grad_(True)
    (ops.oct_minmax_patchify(y3, mean, std, patch).float() * w_out).sum().backward()
    assert torch.equal(y1.grad, y3.grad)


@pytest.mark.gpu
@pytest.mark.parametrize("shape,cout", [((3, 8, 64, 64), 3), ((4, 8, 224, 224), 3), ((2, 16, 40, 52), 3), ((1, 5, 33, 36), 2)])
def test_oct_slice_projection_matches_conv2d(shape, cout):
    """proj_per_3d_slice(image / 255) (trainers/GLP_OT_SVLoRA.py:587-595, :684): own forward and weight-gradient kernels against
    torch's convolution in fp64 on the same inputs (raw 0..255 slices; the input needs no gradient); ragged tiles included."""
    from fairfedmed_b200 import ops
    torch.manual_seed(11)
    bp, cin, h, w = shape
    x = torch.randint(0, 256, shape).float().to("cuda:0")
    wt = (torch.randn(cout, cin, 5, 5) * cin ** -0.5).to("cuda:0").requires_grad_(True)
    b = torch.randn(cout).to("cuda:0").requires_grad_(True)
    dy = torch.randn(bp, cout, h, w).to("cuda:0")
    assert ops.oct_slice_conv_supported(x, wt, 2)
    y = ops.oct_slice_conv(x, wt, b, 1.0 / 255.0)
    y.backward(dy)
    w64, b64 = wt.detach().double().requires_grad_(True), b.detach().double().requires_grad_(True)
    ref = F.conv2d(x.double() / 255.0, w64, b64, padding=2)
    ref.backward(dy.double())
    torch.testing.assert_close(y.detach().double(), ref.detach(), rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(wt.grad.double(), w64.grad, rtol=2e-4, atol=2e-4 * float(w64.grad.abs().max()))
    torch.testing.assert_close(b.grad.double(), b64.grad, rtol=2e-4, atol=2e-4 * float(b64.grad.abs().max()))
    w2, b2 = wt.detach().clone().requires_grad_(True), b.detach().clone().requires_grad_(True)
    ops.oct_slice_conv(x, w2, b2, 1.0 / 255.0).backward(dy)
    assert torch.equal(wt.grad, w2.grad) and torch.equal(b.grad, b2.grad)        # deterministic


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shape,k", [((3, 64, 16, 16), 2), ((2, 256, 56, 56), 2), ((2, 8, 12, 20), 4)])
def test_avgpool_nhwc_matches_torch(shape, k, dtype):
    """nn.AvgPool2d(k) of the ResNet trunk (clip/model.py:30, :42, :108) on channels-last fp32 activations: forward and backward
    against torch."""
    from fairfedmed_b200 import ops
    torch.manual_seed(2)
    x = torch.randn(shape, device="cuda:0").to(dtype).contiguous(memory_format=torch.channels_last)
    assert ops.avgpool_nhwc_supported(x, k)
    x1 = x.clone().requires_grad_(True)
    y = ops.avgpool_nhwc(x1, k)
    dy = torch.randn_like(y)
    y.backward(dy)
    x2 = x.clone().requires_grad_(True)
    ref = F.avg_pool2d(x2, k)
    ref.backward(dy)
    tol = dict(rtol=1e-6, atol=1e-6) if dtype == torch.float32 else dict(rtol=2 ** -7, atol=2 ** -8)
    torch.testing.assert_close(y, ref, **tol)
    torch.testing.assert_close(x1.grad, x2.grad, **tol)
    assert y.is_contiguous(memory_format=torch.channels_last) and x1.grad.is_contiguous(memory_format=torch.channels_last)


@pytest.mark.gpu
def test_widen_bf16_is_exact_and_differentiable():
    """The adapters' bf16 output back into the fp32 ResNet trunk (the `.to(x.dtype)` of trainers/GLP_OT_SVLoRA.py:479-482):
    bit-exact widening, bf16 gradient on the way back, fallback for shapes the vector kernel does not take."""
    from fairfedmed_b200 import ops
    torch.manual_seed(4)
    x = torch.randn(1000, 256, device="cuda:0").bfloat16().requires_grad_(True)
    y = ops.widen_bf16(x)
    assert y.dtype == torch.float32 and torch.equal(y, x.detach().float())
    g = torch.randn_like(y)
    y.backward(g)
    assert x.grad.dtype == torch.bfloat16 and torch.equal(x.grad, g.to(torch.bfloat16))
    odd = torch.randn(7, 3, device="cuda:0").bfloat16()
    assert torch.equal(ops.widen_bf16(odd), odd.float())


@pytest.mark.gpu
@pytest.mark.parametrize("relu", [True, False])
@pytest.mark.parametrize("shape", [(4, 64, 14, 14), (64, 256, 56, 56), (2, 8, 3, 5)])
def test_batchnorm_relu_matches_torch(shape, relu):
    """Training-mode nn.BatchNorm2d (+ ReLU) of the ResNet trunk (clip/model.py:18-58) on channels-last fp32 activations: output,
    running statistics and the gradients of x / gamma / beta against torch in fp64."""
    from fairfedmed_b200 import ops
    torch.manual_seed(9)
    c = shape[1]
    x = (torch.randn(shape, device="cuda:0") * 2.0 + 0.7).contiguous(memory_format=torch.channels_last)
    gamma = (torch.rand(c, device="cuda:0") + 0.5).requires_grad_(True)
    beta = (0.3 * torch.randn(c, device="cuda:0")).requires_grad_(True)
    rm, rv = torch.zeros(c, device="cuda:0"), torch.ones(c, device="cuda:0")
    dy = torch.randn(shape, device="cuda:0").contiguous(memory_format=torch.channels_last)
    x1 = x.clone().requires_grad_(True)
    y = ops.batchnorm_relu(x1, gamma, beta, rm, rv, 0.1, 1e-5, relu)
    y.backward(dy)
    x2 = x.double().requires_grad_(True)
    g2, b2 = gamma.detach().double().requires_grad_(True), beta.detach().double().requires_grad_(True)
    rm2, rv2 = torch.zeros(c, device="cuda:0", dtype=torch.float64), torch.ones(c, device="cuda:0", dtype=torch.float64)
    ref = F.batch_norm(x2, rm2, rv2, g2, b2, True, 0.1, 1e-5)
    if relu:
        ref = F.relu(ref)
    ref.backward(dy.double())
    torch.testing.assert_close(y.detach().double(), ref.detach(), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(rm.double(), rm2, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(rv.double(), rv2, rtol=1e-5, atol=1e-6)
    scale = float(x2.grad.abs().max())
    torch.testing.assert_close(x1.grad.double(), x2.grad, rtol=1e-3, atol=2e-5 * max(scale, 1.0))
    torch.testing.assert_close(gamma.grad.double(), g2.grad, rtol=1e-4, atol=1e-4 * float(g2.grad.abs().max()))
    torch.testing.assert_close(beta.grad.double(), b2.grad, rtol=1e-4, atol=1e-4 * float(b2.grad.abs().max()))


@pytest.mark.gpu
def test_add_relu_matches_torch():
    """relu(out + identity) closing a bottleneck (clip/model.py:56-58), channels-last fp32, forward and both gradients."""
    from fairfedmed_b200 import ops
    torch.manual_seed(6)
    a = torch.randn(4, 64, 14, 14, device="cuda:0").contiguous(memory_format=torch.channels_last)
    b = torch.randn(4, 64, 14, 14, device="cuda:0").contiguous(memory_format=torch.channels_last)
    a1, b1 = a.clone().requires_grad_(True), b.clone().requires_grad_(True)
    y = ops.add_relu(a1, b1)
    dy = torch.randn_like(y)
    y.backward(dy)
    a2, b2 = a.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ref = torch.relu(a2 + b2)
    ref.backward(dy)
    assert torch.equal(y, ref) and torch.equal(a1.grad, a2.grad) and torch.equal(b1.grad, b2.grad)
    assert y.is_contiguous(memory_format=torch.channels_last)
