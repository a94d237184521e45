"""GLP_OT_SVLoRA trainer — the plug-in boundary of the reference (trainers/GLP_OT_SVLoRA.py:767-1053 on top of
Dassl's TrainerX, Dassl/dassl/engine/trainer.py:345-741), registered under the same TRAINER_REGISTRY name.

What federated_main.py touches is kept: ctor `(cfg)`, `.model` (+ state_dict / load_state_dict), `.dm.dataset.
classnames`, `.fed_train_loader_x_dict[i].dataset` (`len`, `count_by_attribute`), `.fed_before_train()`,
`.train(idx=, global_epoch=, is_fed=, is_last_client=)`, `.test(idx=, current_epoch=)` -> list whose [0..3] are
acc / err / macro-F1 / AUC, `.fed_after_train()`.

B200-first differences (documented in DESIGN.md):
  * all trainable tensors (prompt ctx, lora_A / lora_S / lora_B, OCT slice projection) are VIEWS into one flat fp32
    buffer; gradients likewise.  The optimizer is ONE fused kernel over that buffer which also reproduces the
    reference's double optimizer.step() per iteration (same SGD object registered under two model names,
    :864-871) and the doubly-stepped StepLR; federated aggregation all-reduces the very same buffer;
  * batches arrive in pinned memory and are copied asynchronously; loss / accuracy stay on the device and are
    only synchronised when a caller asks for python floats;
  * evaluation accumulates probabilities on the device and computes every fairness metric from one
    segmented-sort kernel pass (fairfedmed_b200/metrics.py).
"""
from __future__ import annotations

import os
import time
from collections import OrderedDict
from typing import Optional

import numpy as np
import torch
import torch.nn.functional as F

from . import clip_weights
from . import metrics as M
from . import ops
from .clip_model import CustomCLIP
from .config import ATTRIBUTE_GROUPS
from .data import SyntheticDataManager
from .fed_utils import FlatSpec, build_spec
from .modules import apply_lora_to_model
from .registry import TRAINER_REGISTRY


def _cfg_get(cfg, path: str, default=None):
    """cfg.A.B.C or `default` when any level is missing (the reference's yacs cfg has no MODEL_ARCH / SYNTHETIC)."""
    node = cfg
    for part in path.split("."):
        try:
            node = getattr(node, part)
        except (AttributeError, KeyError):
            return default
    return node


@TRAINER_REGISTRY.register()
class GLP_OT_SVLoRA:
    # Integration hooks (class attributes so that Dassl's `build_trainer(cfg)`, which only passes cfg, can be served):
    #   data_manager_factory(cfg) -> object with fed_train_loader_x_dict / fed_test_loader_x_dict / dataset.classnames
    #   clip_loader(cfg)          -> pretrained CLIP state dict (or module / TorchScript archive with .state_dict())
    #   tokenize(str)             -> int tensor [1, 77] (clip.tokenize)
    data_manager_factory = None
    clip_loader = None
    tokenize = None
    _require_cuda = True              # construction-only escape for CPU-side surface tests; training always needs CUDA

    def __init__(self, cfg, device: Optional[torch.device] = None, data_manager=None, clip_state_dict=None):
        if self._require_cuda and not torch.cuda.is_available():
            raise RuntimeError("GLP_OT_SVLoRA (fairfedmed_b200) needs a CUDA device: the training path is made of "
                               "sm_100a kernels and has no CPU fallback")
        self.check_cfg(cfg)
        self.cfg = cfg
        self.device = device or (torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available()
                                 else torch.device("cpu"))
        self.synthetic = bool(_cfg_get(cfg, "DATASET.SYNTHETIC", False))
        self.build_data_loader(data_manager)
        self.start_epoch = self.epoch = 0
        self.max_epoch = cfg.OPTIM.MAX_EPOCH
        self.output_dir = _cfg_get(cfg, "OUTPUT_DIR", "output")
        # "epoch": loss / acc / AUC of every step stay on the device and are read back once per epoch (the step is
        # replayed from a CUDA graph at the benchmarked rate); "step": python floats after every step like the
        # reference (trainers/GLP_OT_SVLoRA.py:952-970) — one host sync per step.
        self.step_metrics = "epoch"
        self.sync_metrics = True          # forward_backward() returns python floats (reference behaviour)
        self.step_auc = True              # per-step training AUC like the reference (:964-970)
        self.use_cuda_graph = os.environ.get("FFM_TRAIN_GRAPH", "1") != "0"
        self.sched_steps = 0              # StepLR.step() calls so far
        self.first_step = True
        self._graph = None
        self._graph_key = None
        self.build_model(clip_state_dict)

    # ------------------------------------------------------------------ configuration
    def check_cfg(self, cfg):
        assert cfg.TRAINER.GLP_OT.PREC in ["bf16", "fp16", "fp32", "amp"]

    def retrieval_attributes(self, attr_name):
        try:
            return ATTRIBUTE_GROUPS[self.cfg.DATASET.NAME][attr_name]
        except KeyError as e:
            raise NotImplementedError(f"{self.cfg.DATASET.NAME}/{attr_name}") from e

    @property
    def num_groups(self) -> int:
        if self.cfg.TRAINER.GLP_OT_LORA.DISABLE_ATTR:
            return 1
        return len(self.retrieval_attributes(self.cfg.DATASET.ATTRIBUTE_TYPE))

    # ------------------------------------------------------------------ data
    def build_data_loader(self, data_manager=None):
        """SimpleTrainer.build_data_loader (Dassl/dassl/engine/trainer.py:382-400).  Resolution order: the explicit
        argument, the class-level `data_manager_factory`, the synthetic FairFedMed-shaped manager when
        `cfg.DATASET.SYNTHETIC` is set, the reference's own `Dassl.dassl.data.DataManager` when it is importable.
        Anything else is an error: silently training on random data would be worse than failing."""
        cfg = self.cfg
        dm = data_manager
        if dm is None and type(self).data_manager_factory is not None:
            dm = type(self).data_manager_factory(cfg)
        if dm is None and self.synthetic:
            dm = SyntheticDataManager(cfg)
        if dm is None:
            try:
                from Dassl.dassl.data import DataManager          # the reference tree on sys.path
            except Exception as e:
                raise RuntimeError(
                    "GLP_OT_SVLoRA: no DataManager — pass data_manager=, set GLP_OT_SVLoRA.data_manager_factory, put the "
                    "reference's Dassl on sys.path, or set cfg.DATASET.SYNTHETIC = True for synthetic batches") from e
            dm = DataManager(cfg)
        self.dm = dm
        self.fed_train_loader_x_dict = dm.fed_train_loader_x_dict
        self.fed_test_loader_x_dict = dm.fed_test_loader_x_dict
        self.classnames = list(getattr(dm, "classnames", None) or dm.dataset.classnames)
        self.lab2cname = getattr(dm, "lab2cname", None) or getattr(dm.dataset, "lab2cname", None)
        self.num_classes = getattr(dm, "num_classes", len(self.classnames))

    # ------------------------------------------------------------------ model / optimizer
    def _arch(self, clip_sd):
        """Tower shapes: explicit cfg.MODEL_ARCH (tests, shrunken towers) > the checkpoint's own shapes >
        cfg.MODEL.BACKBONE.NAME (the only thing the reference's cfg holds, federated_main.py:46-47)."""
        arch = _cfg_get(self.cfg, "MODEL_ARCH")
        name = self.cfg.MODEL.BACKBONE.NAME
        if arch is not None:
            a = dict(arch)
            if name.startswith("RN") and not isinstance(a["VISION_LAYERS"], (tuple, list)):
                a.update(clip_weights.ARCH_BY_BACKBONE[name])           # CLIP ResNet: (3, 4, 6, 3), width 64, embed 1024
            return a
        if clip_sd is not None:
            return clip_weights.arch_from_clip_state_dict(clip_sd)
        return clip_weights.arch_for_backbone(name)

    def build_model(self, clip_state_dict=None):
        cfg = self.cfg
        ot = cfg.TRAINER.GLP_OT
        if cfg.TRAINER.GLP_OT.PREC == "fp32":
            raise NotImplementedError("the B200 path computes in bf16 with fp32 accumulation; use the oracle for fp32")
        if clip_state_dict is None and type(self).clip_loader is not None:
            clip_state_dict = type(self).clip_loader(cfg)
        clip_sd = clip_weights.resolve_state_dict(clip_state_dict)
        if clip_sd is None and not self.synthetic:
            raise RuntimeError(
                "GLP_OT_SVLoRA: no pretrained CLIP weights — pass clip_state_dict= or set GLP_OT_SVLoRA.clip_loader "
                "(e.g. lambda cfg: load_clip_to_cpu(cfg).state_dict()); random-init towers are only built when "
                "cfg.DATASET.SYNTHETIC = True")
        arch = self._arch(clip_sd)
        seed = cfg.SEED if cfg.SEED is not None and cfg.SEED >= 0 else 0
        torch.manual_seed(seed)
        is_3d = cfg.DATASET.MODALITY_TYPE in {"oct_bscans", "oct_bscans_3d", "mac_onh", "onh_mac"}
        prompt_buffers = None
        if clip_sd is not None:
            if type(self).tokenize is None:
                raise RuntimeError("GLP_OT_SVLoRA: pretrained weights need GLP_OT_SVLoRA.tokenize (clip.tokenize) to build "
                                   "the prompt token buffers")
            tok = clip_weights.tokenize_prompts(self.classnames, ot.N_CTX, type(self).tokenize)
            prompt_buffers = clip_weights.prompt_buffers_from_clip(clip_sd["token_embedding.weight"], tok, ot.N,
                                                                   ot.N_CTX)
        self.model = CustomCLIP(
            classnames=self.classnames, n_prompts=ot.N, n_ctx=ot.N_CTX, ot=ot.OT, eps=ot.EPS,
            thresh=ot.THRESH, max_iter=ot.MAX_ITER, top_percent=ot.TOP_PERCENT, image_resolution=cfg.INPUT.SIZE[0],
            vision_layers=arch["VISION_LAYERS"], vision_width=arch["VISION_WIDTH"],
            vision_patch_size=arch["PATCH"] or 16, embed_dim=arch["EMBED"], text_width=arch["TEXT_WIDTH"],
            text_layers=arch["TEXT_LAYERS"], text_heads=arch["TEXT_HEADS"], context_length=arch["CONTEXT"],
            dim_per_3d_slice=cfg.DATASET.DIM_PER_3D_SLICE if is_3d else None, dataset=cfg.DATASET.NAME, seed=seed,
            prompt_buffers=prompt_buffers)
        if clip_sd is not None:
            clip_weights.load_clip_into(self.model, clip_sd)
        # freeze everything but the prompt learner / OCT projection / BatchNorm2d affine parameters of the ResNet trunk
        # (:822-829), then wrap the MLP linears or the 1x1 convolutions (:834-842)
        bn_params = {id(p) for m_ in self.model.modules() if isinstance(m_, torch.nn.BatchNorm2d)
                     for p in m_.parameters()}
        for name, p in self.model.named_parameters():
            p.requires_grad_("prompt_learner" in name or "proj_per_3d_slice" in name or id(p) in bn_params)
        init_w = _cfg_get(cfg, "MODEL.INIT_WEIGHTS", "")
        if init_w:                                    # load_pretrained_weights(self.model.prompt_learner, ...) :831-832
            ck = torch.load(init_w, map_location="cpu")
            ck = ck.get("state_dict", ck)
            ck = {k: v for k, v in ck.items() if "token_prefix" not in k and "token_suffix" not in k}
            self.model.prompt_learner.load_state_dict(ck, strict=False)
        lora = cfg.TRAINER.GLP_OT_LORA
        apply_lora_to_model(self.model, lora.UNFREEZE_IMAGE_ENCODER, rank=lora.RANK, alpha=lora.ALPHA,
                            lora_type=lora.TYPE, global_s=lora.GLOBAL_S, num_attrs=self.num_groups,
                            adapt_attention=bool(_cfg_get(cfg, "TRAINER.GLP_OT_LORA.ADAPT_ATTENTION", False)))
        self.model.to(self.device)
        self._flatten_trainables()
        self.base_lr = cfg.OPTIM.LR
        self.lr_dev = torch.full((1,), float(self.current_lr()), device=self.device, dtype=torch.float32)

    def _flatten_trainables(self):
        """ONE flat fp32 buffer holds (a) every trainable tensor and (b) every floating-point buffer that training
        changes — the BatchNorm running statistics of the ResNet trunk — because the reference averages the whole
        state dict each round (utils/fed_utils.py:63-98, incl. running_mean / running_var / num_batches_tracked).
        Parameters and running statistics are re-pointed to views of it, so the optimizer (first `n_trainable`
        elements), the aggregation and the per-client reset all work on the buffer in place."""
        named = [(n, p) for n, p in self.model.named_parameters() if p.requires_grad]
        # tensors whose size is not a multiple of 4 floats (the 3-element bias of the OCT slice projection) go last, so
        # every other tensor starts on a 16-byte boundary (library BatchNorm kernels and the float4 paths of this
        # package assume that of a tensor's first element)
        named = [x for x in named if x[1].numel() % 4 == 0] + [x for x in named if x[1].numel() % 4 != 0]
        stats = []                                     # (key, module, buffer name)
        bns = [(mname, mod) for mname, mod in self.model.named_modules()
               if isinstance(mod, torch.nn.modules.batchnorm._BatchNorm) and mod.track_running_stats]
        for bname in ("running_mean", "running_var"):
            for mname, mod in bns:
                stats.append((f"{mname}.{bname}", mod, bname))
        for mname, mod in bns:                         # one-element int64 batch counters (as floats) at the very end
            stats.append((f"{mname}.num_batches_tracked", mod, "num_batches_tracked"))
        n_train = sum(p.numel() for _, p in named)
        n_pad = (-n_train) % 4 if stats else 0         # keeps the running statistics 16-byte aligned
        n_stats = sum(getattr(m_, b).numel() for _, m_, b in stats)
        self.n_trainable = n_train
        self.flat_all = torch.zeros(n_train + n_pad + n_stats, device=self.device, dtype=torch.float32)
        self.flat_params = self.flat_all[:n_train]
        self.flat_grads = torch.zeros(n_train, device=self.device, dtype=torch.float32)
        self.flat_mom = torch.zeros(n_train, device=self.device, dtype=torch.float32)
        off = 0
        for name, p in named:
            n = p.numel()
            self.flat_params[off:off + n].copy_(p.data.reshape(-1))
            p.data = self.flat_params[off:off + n].view(p.shape)
            p.grad = self.flat_grads[off:off + n].view(p.shape)
            if any(k in name for k in ("lora_A", "lora_B", "lora_S")) and "attnpool" not in name:
                # the fused backward kernels write these gradients straight into the flat buffer (ops._direct_grad):
                # one backward per zero_grad, which is how every step of this trainer runs.  The attention-pool
                # LoRA of the ResNet goes through merged weights and plain autograd, so it keeps accumulating.
                p._ffm_direct_grad = p.grad
            off += n
        sd = {n: p for n, p in named}
        keys = [n for n, _ in named]
        if n_pad:
            sd["__pad__"] = self.flat_all[off:off + n_pad]
            keys.append("__pad__")
            off += n_pad
        self._int_stats = []                           # (module, flat view): int64 counters mirrored as floats
        for key, mod, bname in stats:
            buf = getattr(mod, bname)
            n = buf.numel()
            view = self.flat_all[off:off + n].view(buf.shape)
            view.copy_(buf)
            if buf.is_floating_point():
                mod._buffers[bname] = view             # BatchNorm updates the flat buffer in place from now on
            else:
                self._int_stats.append((mod, bname, view))
            sd[key] = view
            keys.append(key)
            off += n
        self.flat_spec: FlatSpec = build_spec(sd, keys=keys,
                                              num_groups=None if self.num_groups == 1 else self.num_groups)
        self.trainable_names = [n for n, _ in named]
        self.aggregated_names = keys

    def current_lr(self) -> float:
        """StepLR(step_size, gamma) after `sched_steps` calls of .step() (Dassl/dassl/optim/lr_scheduler.py:83+)."""
        step = self.cfg.OPTIM.STEPSIZE[0] if isinstance(self.cfg.OPTIM.STEPSIZE, (tuple, list)) else self.cfg.OPTIM.STEPSIZE
        if step <= 0:
            step = self.max_epoch
        return self.base_lr * (self.cfg.OPTIM.GAMMA ** (self.sched_steps // max(step, 1)))

    get_current_lr = current_lr

    def _n_opt_steps(self) -> int:
        return 2 if self.cfg.TRAINER.GLP_OT_LORA.UNFREEZE_IMAGE_ENCODER else 1

    def model_backward_and_update(self, loss):
        """zero_grad x2, backward, optimizer.step() x2 (engine/trainer.py:333-342 with the shared optimizer, F6)."""
        self.flat_grads.zero_()
        if self.sync_metrics and not bool(torch.isfinite(loss.detach())):   # detect_anomaly (engine/trainer.py:260-262)
            raise FloatingPointError("Loss is infinite or NaN!")
        loss.backward()
        ops.join_direct_grad_writes()
        o = self.cfg.OPTIM
        ops.sgd_step_(self.flat_params, self.flat_grads, self.flat_mom, self.current_lr(), o.MOMENTUM, o.WEIGHT_DECAY,
                      self._n_opt_steps(), self.first_step)
        self.first_step = False

    def update_lr(self):
        self.sched_steps += self._n_opt_steps()
        self.lr_dev.fill_(float(self.current_lr()))

    # ------------------------------------------------------------------ batches
    def _parse(self, batch):
        image = batch["img"].to(self.device, non_blocking=True)
        label = batch["label"].to(self.device, non_blocking=True)
        attrs = batch["attrs"].t()
        idx = self.cfg.DATASET.ATTRIBUTES.index(self.cfg.DATASET.ATTRIBUTE_TYPE)
        tgt = None if self.cfg.TRAINER.GLP_OT_LORA.DISABLE_ATTR else attrs[idx].contiguous()
        return image, label, attrs, tgt

    parse_batch_train = _parse
    parse_batch_test = _parse

    # ------------------------------------------------------------------ one SGD step (HOT LOOP body)
    def _loss_and_metrics(self, output, label, attr):
        """Cross-entropy (+ the detached fairness term, which only changes the reported VALUE, :930-948), accuracy and
        the softmax probabilities — all device tensors, no host sync."""
        cls_loss = F.cross_entropy(output, label)
        loss = cls_loss
        lam = self.cfg.TRAINER.LAMBDA_FAIRNESS
        if attr is not None and lam != 0.0:
            with torch.no_grad():
                probs = F.softmax(output, dim=1)
                correct = probs[torch.arange(len(label), device=label.device), label]
                a = attr.to(self.device)
                # mean over the groups PRESENT in the batch of |conf_g - mean_g conf_g|, without the data-dependent
                # shapes of torch.unique (graph-capturable): absent groups get weight zero
                G = max(self.num_groups, 1)
                onehot = F.one_hot(a.long(), G).to(correct.dtype)                       # [B, G]
                cnt = onehot.sum(0)
                present = (cnt > 0).to(correct.dtype)
                conf = 1 - (onehot * correct[:, None]).sum(0) / cnt.clamp_min(1)
                mean_conf = (conf * present).sum() / present.sum()
                fairness = ((conf - mean_conf).abs() * present).sum() / present.sum()
            loss = cls_loss + lam * fairness
        return cls_loss, loss

    def forward_backward(self, batch, is_last_client=False):
        """The eager step (one launch per kernel).  `run_epoch` replays the same step from a CUDA graph."""
        image, label, _, attr = self.parse_batch_train(batch)
        output = self.model(image, attr)
        if output is None:
            raise FloatingPointError("transport plan contains NaN (CustomCLIP.forward returned None)")
        cls_loss, loss = self._loss_and_metrics(output, label, attr)
        self.model_backward_and_update(loss)
        with torch.no_grad():
            acc = (output.argmax(dim=1) == label).float().mean() * 100.0
        summary = {"loss": loss.detach(), "acc": acc}
        if self.step_auc:
            prob = output.detach().softmax(-1)
            summary["auc"] = M.compute_auc(prob, label)
        if self.sync_metrics:
            summary = {k: (v.item() if torch.is_tensor(v) else float(v)) for k, v in summary.items()}
        if (self.batch_idx + 1) == self.num_batches:
            self.update_lr()
        return summary

    # ------------------------------------------------------------------ CUDA-graph replay of the step
    def capture_step_graph(self, example_batch, warmup: int = 3):
        """Capture forward + backward + fused double-SGD of ONE step into a CUDA graph (launch-bound otherwise:
        ~1000 kernel launches per step).  Inputs live in static device buffers; `forward_backward_graphed` copies a
        batch in and replays.  The learning rate is read from device memory (`lr_dev`, ffm_sgd_step_dev_lr), so the
        graph survives StepLR changes; the momentum buffer starts at zero, which makes torch's lazily initialised
        first step the same arithmetic.  Warm-up steps run for real (library handles, allocator pools) and are rolled
        back: parameters, momentum, BatchNorm statistics and RNG-free state are restored afterwards."""
        self.model.train()
        image, label, _, attr = self.parse_batch_train(example_batch)
        self._g_img = image.clone()
        self._g_label = label.clone()
        self._g_attr = None if attr is None else attr.to(self.device).clone()
        o = self.cfg.OPTIM
        n_steps = self._n_opt_steps()
        prev_check = self.model.check_nan
        self.model.check_nan = False              # the NaN flag of the plan is read back by the caller, not in-graph

        def body():
            output = self.model(self._g_img, self._g_attr)
            _, loss = self._loss_and_metrics(output, self._g_label, self._g_attr)
            self.flat_grads.zero_()
            loss.backward()
            ops.join_direct_grad_writes()
            ops.sgd_step_dev_lr_(self.flat_params, self.flat_grads, self.flat_mom, self.lr_dev, o.MOMENTUM,
                                 o.WEIGHT_DECAY, n_steps)
            with torch.no_grad():
                acc = (output.argmax(dim=1) == self._g_label).float().mean() * 100.0
                prob = output.detach().float().softmax(-1)
            return loss.detach(), acc, prob

        # The step's main stream is captured at HIGH priority; the side streams the model forks (text tower, adapter
        # preparation, parameter gradients) keep the default, lower one.  Kernel nodes inherit the priority of the
        # stream they were captured on, so whenever a main-chain kernel and side work are both ready the block
        # scheduler serves the critical path first and the side work fills what is left.
        prio = -1 if os.environ.get("FFM_GRAPH_PRIO", "1") != "0" else 0
        side = torch.cuda.Stream(device=self.device, priority=prio)
        snap_all, snap_mom = self.flat_all.clone(), self.flat_mom.clone()
        snap_int = [getattr(m_, b).clone() for m_, b, _ in self._int_stats]
        try:
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(warmup):
                    body()
            torch.cuda.current_stream().wait_stream(side)
            self._graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph, stream=side):
                self._g_loss, self._g_acc, self._g_prob = body()
        finally:
            self.model.check_nan = prev_check
            self.flat_all.copy_(snap_all)
            self.flat_mom.copy_(snap_mom)
            for (m_, b, _), v in zip(self._int_stats, snap_int):
                getattr(m_, b).copy_(v)
        self._graph_key = self._batch_key(example_batch)
        return self._graph

    def _batch_key(self, batch):
        return (tuple(batch["img"].shape), batch["img"].dtype, tuple(batch["label"].shape),
                tuple(batch["attrs"].shape), bool(self.cfg.TRAINER.GLP_OT_LORA.DISABLE_ATTR))

    def forward_backward_graphed(self, batch):
        """Replay the captured step on `batch` (device or pinned-host tensors). Returns device tensors."""
        self._g_img.copy_(batch["img"], non_blocking=True)
        self._g_label.copy_(batch["label"], non_blocking=True)
        if self._g_attr is not None:
            idx = self.cfg.DATASET.ATTRIBUTES.index(self.cfg.DATASET.ATTRIBUTE_TYPE)
            self._g_attr.copy_(batch["attrs"][:, idx], non_blocking=True)
        self._graph.replay()
        return {"loss": self._g_loss, "acc": self._g_acc, "prob": self._g_prob}

    # ------------------------------------------------------------------ epoch / federated hooks
    def fed_before_train(self, is_global=False):
        self.start_epoch = 0
        self.total_time_start = self.time_start = time.time()

    def fed_after_train(self):
        pass

    def before_train(self, is_fed=False):
        self.start_epoch = 0
        self.time_start = time.time()

    def before_epoch(self):
        pass

    def after_epoch(self, idx=-1, global_epoch=-1):
        """SimpleTrainer.after_epoch (Dassl/dassl/engine/trainer.py:497-521): checkpoint every TRAIN.CHECKPOINT_FREQ
        local epochs and after the last one, `epoch{g}[_client{idx}].pth` under OUTPUT_DIR."""
        freq = _cfg_get(self.cfg, "TRAIN.CHECKPOINT_FREQ", 0) or 0
        last_epoch = (self.epoch + 1) == self.max_epoch
        meet = freq > 0 and (self.epoch + 1) % freq == 0
        if not (meet or last_epoch) or not _cfg_get(self.cfg, "TRAIN.SAVE_CHECKPOINTS", freq > 0):
            return None
        os.makedirs(self.output_dir, exist_ok=True)
        name = f"epoch{global_epoch}.pth" if idx == -1 else f"epoch{global_epoch}_client{idx}.pth"
        filename = os.path.join(self.output_dir, name)
        self.save_model_with_grad(filename)
        return filename

    def after_train(self, idx=-1, epoch=0, is_fed=False):
        if not _cfg_get(self.cfg, "TEST.NO_TEST", False) and not is_fed:
            self.test(idx=idx, current_epoch=epoch)

    def set_model_mode(self, mode="train", names=None):
        self.model.train() if mode == "train" else self.model.eval()

    def _staged_batches(self, loader):
        """Iterate `loader` with the NEXT batch's host->device copy in flight on a copy stream while the current step
        runs (two rotating device slots, no allocation per step)."""
        copy_stream = self.__dict__.setdefault("_copy_stream", torch.cuda.Stream(device=self.device))
        slots, free = [None, None], [None, None]
        it = iter(loader)

        def stage(i):
            try:
                host = next(it)
            except StopIteration:
                return None
            sl = i % 2
            key = self._batch_key(host)
            if slots[sl] is None or slots[sl][0] != key:
                slots[sl] = (key, {k: torch.empty_like(v, device=self.device) for k, v in host.items()
                                   if torch.is_tensor(v)})
                free[sl] = None
            with torch.cuda.stream(copy_stream):
                if free[sl] is not None:
                    copy_stream.wait_event(free[sl])
                for k, v in slots[sl][1].items():
                    v.copy_(host[k], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            return slots[sl][1], ev, sl, host

        nxt = stage(0)
        i = 0
        while nxt is not None:
            dev_batch, ev, sl, host = nxt
            nxt = stage(i + 1)
            torch.cuda.current_stream().wait_event(ev)
            yield dev_batch
            done = torch.cuda.Event()
            done.record()
            free[sl] = done
            i += 1

    def run_epoch(self, idx=-1, global_epoch=0, is_fed=True, is_last_client=False):
        """TrainerX.run_epoch (Dassl/dassl/engine/trainer.py:685-741).  With `step_metrics == "epoch"` (default) every
        step is a replay of the captured CUDA graph with its batch staged ahead on a copy stream, loss / acc / probs
        stay on the device and are read back once at the end; `"step"` is the reference's behaviour (python floats,
        per-step AUC, one host sync per step) on the eager path."""
        self.model.train()
        loader = self.fed_train_loader_x_dict[idx]
        self.num_batches = len(loader)
        deferred = self.step_metrics == "epoch" and self.use_cuda_graph
        if deferred and self._graph is None:
            try:                                           # capture from the first staged batch of this loader
                first = next(iter(loader), None)
                if first is not None:
                    self.capture_step_graph({k: v.to(self.device) for k, v in first.items() if torch.is_tensor(v)})
            except Exception as e:                         # capture is an optimisation: fall back to eager launches
                import warnings
                warnings.warn(f"CUDA-graph capture of the training step failed ({type(e).__name__}: {e}); "
                              "running eager launches")
                torch.cuda.synchronize()
                self._graph, self.use_cuda_graph, deferred = None, False, False
        if not deferred:
            last = None
            for self.batch_idx, batch in enumerate(loader):
                last = self.forward_backward(batch, is_last_client)
            self.last_epoch_summaries = None
            return last
        n = self.num_batches
        hist = torch.zeros((max(n, 1), 2), device=self.device, dtype=torch.float32)
        probs, labels = [], []
        for self.batch_idx, batch in enumerate(self._staged_batches(loader)):
            if self._graph is None or self._graph_key != self._batch_key(batch):
                self.capture_step_graph(batch)
            out = self.forward_backward_graphed(batch)
            hist[self.batch_idx, 0].copy_(out["loss"])
            hist[self.batch_idx, 1].copy_(out["acc"])
            if self.step_auc:
                probs.append(out["prob"].clone())
                labels.append(self._g_label.clone())
            if (self.batch_idx + 1) == self.num_batches:
                self.update_lr()
        host = hist.cpu()                                      # the one synchronisation of the epoch
        if self.model.OT != "None" and self.model.last_status is not None and int(self.model.last_status[1].item()):
            raise FloatingPointError("transport plan contains NaN (CustomCLIP.forward returned None)")
        if not bool(torch.isfinite(host[:, 0]).all()):
            raise FloatingPointError("Loss is infinite or NaN!")
        summaries = [{"loss": float(host[i, 0]), "acc": float(host[i, 1])} for i in range(n)]
        if self.step_auc and n:
            # every step's training AUC from ONE pass of the segmented-sort kernel: the step index is the "group"
            bsz = probs[0].shape[0]
            for lo in range(0, n, 254):
                hi = min(n, lo + 254)
                ids = torch.arange(hi - lo, device=self.device, dtype=torch.int32).repeat_interleave(bsz)[None, :]
                counts = M.group_counts(torch.cat(probs[lo:hi]), torch.cat(labels[lo:hi]), ids, max_groups=hi - lo)
                for i in range(lo, hi):
                    try:
                        summaries[i]["auc"] = M._auc_from_row(counts.slot(0, i - lo))
                    except ValueError:                         # single label in the batch (:960-962)
                        summaries[i]["auc"] = 1
        self.last_epoch_summaries = summaries
        return summaries[-1] if summaries else None

    def train(self, idx=-1, global_epoch=0, is_fed=False, is_last_client=False, **_unused):
        """TrainerBase.train (Dassl/dassl/engine/trainer.py:283-294): before_train, max_epoch x (before_epoch, run_epoch,
        after_epoch), after_train."""
        last = None
        self.before_train(is_fed)
        for self.epoch in range(self.start_epoch, self.max_epoch):
            self.before_epoch()
            last = self.run_epoch(idx, global_epoch, is_fed, is_last_client)
            self.after_epoch(idx, global_epoch)
        self.after_train(idx, global_epoch, is_fed)
        return last

    @torch.no_grad()
    def model_inference(self, image, attr=None):
        return self.model(image, attr)

    @torch.no_grad()
    def test(self, idx=-1, current_epoch=0, split=None):
        """Classification_oph.process / evaluate (evaluation/evaluator_oph.py:37-151) with device-side accumulation."""
        self.model.eval()
        probs, labels, attrs_all = [], [], []
        correct = torch.zeros((), device=self.device)
        total = 0
        for batch in self.fed_test_loader_x_dict[idx]:
            image, label, attrs, tgt = self.parse_batch_test(batch)
            out = self.model_inference(image, tgt).float()
            probs.append(out.softmax(-1))
            labels.append(label)
            attrs_all.append(attrs.to(self.device, non_blocking=True))
            correct += (out.argmax(1) == label).sum()
            total += label.shape[0]
        prob = torch.cat(probs)
        gt = torch.cat(labels)
        attr = torch.cat(attrs_all, dim=1)
        res = M.evalute_comprehensive_perf_scores(prob, gt, attr)
        counts = M.group_counts(prob, gt).slot()
        tp, fp, tn, fn = (float(counts[i]) for i in (M.TP, M.FP, M.TN, M.FN))
        f1s = []
        if tp + fn > 0:
            f1s.append(2 * tp / max(2 * tp + fp + fn, 1e-30) if (2 * tp + fp + fn) else 0.0)
        if tn + fp > 0:
            f1s.append(2 * tn / max(2 * tn + fn + fp, 1e-30) if (2 * tn + fn + fp) else 0.0)
        acc = 100.0 * float(correct.item()) / total
        results = OrderedDict()
        results["accuracy"] = acc
        results["error_rate"] = 100.0 - acc
        results["macro_f1"] = 100.0 * float(np.mean(f1s))
        results["auc"] = 100.0 * res[2]
        (results["overall_acc"], results["esaccs_by_attrs"], results["overall_auc"], results["esaucs_by_attrs"],
         results["aucs_by_attrs"], results["dpds"], results["eods"], results["aods"],
         results["between_group_disparity"]) = res
        self.last_results = results
        return list(results.values())

    # ------------------------------------------------------------------ checkpoints (scope row f4)
    def save_model_with_grad(self, filename):
        """TrainerBase.save_model_with_grad (Dassl/dassl/engine/trainer.py:177-185): trainable parameters + buffers
        under the reference's state-dict keys — the file `federated_main.py` / the evaluation scripts expect."""
        state = {n: p.detach().cpu().clone() for n, p in self.model.named_parameters() if p.requires_grad}
        state.update({n: b.detach().cpu().clone() for n, b in self.model.named_buffers()})
        torch.save(state, filename)

    def load_model_with_grad(self, filename):
        """Inverse of save_model_with_grad: copies INTO the existing parameters (they are views of the flat buffer the
        optimizer and the aggregation work on; parameter objects must persist)."""
        state = torch.load(filename, map_location="cpu")
        missing, unexpected = self.model.load_state_dict(state, strict=False)
        bad = [k for k in missing if k in self.trainable_names]
        buffers = dict(self.model.named_buffers())        # non-persistent buffers are saved (as upstream) but constant
        unexpected = [k for k in unexpected if k not in buffers]
        if bad or unexpected:
            raise KeyError(f"checkpoint does not match the model: missing trainable {bad}, unexpected {unexpected}")
        return self

    def save_flat(self, filename):
        """Adapter-only wire format: the flat fp32 buffer exactly as the per-round all-reduce sends it, plus the key /
        shape table that maps it back to state-dict entries (4.4 MB for ViT-B/16 instead of the reference's 500 MB
        `global_client{idx}_final.pth`, federated_main.py:775-778)."""
        torch.save({"flat": self.get_flat().detach().cpu().clone(), "keys": list(self.flat_spec.keys),
                    "shapes": [tuple(sh) for sh in self.flat_spec.shapes]}, filename)

    def load_flat(self, filename):
        blob = torch.load(filename, map_location="cpu")
        if list(blob["keys"]) != list(self.flat_spec.keys) or blob["flat"].numel() != self.flat_all.numel():
            raise KeyError("flat checkpoint does not match this trainer's trainable tensors")
        self.set_flat(blob["flat"].to(self.device))
        return self

    # ------------------------------------------------------------------ flat-buffer access for aggregation
    def get_flat(self) -> torch.Tensor:
        """Everything the round aggregates: trainable tensors, then BatchNorm running statistics (in place), then the
        float mirrors of BatchNorm's int64 batch counters (refreshed here)."""
        for mod, bname, view in self._int_stats:
            view.copy_(getattr(mod, bname))
        return self.flat_all

    def set_flat(self, flat: torch.Tensor) -> None:
        self.flat_all.copy_(flat)
        for mod, bname, view in self._int_stats:
            getattr(mod, bname).copy_(view)            # float -> int64 truncation, like load_state_dict's copy_
