"""Server aggregation on flat device buffers — drop-in for utils/fed_utils.py of the reference.

`average_weights_EMA` / `average_weights` keep the reference signatures (utils/fed_utils.py:6-100) and operate on
state dicts held by ONE process (the reference's layout: a python list of per-client dicts).
`FederatedAggregator` is the B200 layout: one simulated site per rank / GPU, the weighted sum is ONE NCCL
all-reduce over NVLink of the pre-scaled flat adapter buffer; the tiny count exchange that fixes the weights
is a second (1+G)-float all-reduce.  Both paths share the same two CUDA kernels (csrc/fedavg.cu).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import torch
import torch.distributed as dist

from . import ops


@dataclass
class FlatSpec:
    """Layout of a flat fp32 buffer holding the aggregated tensors, in a fixed key order."""
    keys: List[str]
    shapes: List[torch.Size]
    offsets: List[int]
    kinds: List[int]          # 1: lora_S [G, r] rows weighted per group; 0: scalar client weight
    numel: int
    G: int
    r: int

    def segment_tensors(self, device):
        cached = getattr(self, "_seg", None)
        if cached is None or cached[0] != device:
            kind = torch.tensor(self.kinds, dtype=torch.int32, device=device)
            off = torch.tensor(self.offsets, dtype=torch.int64, device=device)
            ln = torch.tensor([int(torch.Size(s).numel()) for s in self.shapes], dtype=torch.int64, device=device)
            self._seg = (device, kind, off, ln)
            cached = self._seg
        return cached[1], cached[2], cached[3]


def build_spec(reference: Dict[str, torch.Tensor], keys: Optional[Sequence[str]] = None,
               num_groups: Optional[int] = None) -> FlatSpec:
    """Describe `keys` of `reference` (default: every tensor, in dict order; integer tensors such as BatchNorm's
    `num_batches_tracked` travel as floats, which is what the reference's `w * freq` arithmetic turns them into).

    A key is weighted per group when it contains 'lora_S', group counts are supplied and its leading dimension
    equals the number of groups (utils/fed_utils.py:77) — restricted here to 2-D tensors, the only case in which
    the reference's `[:, None]` broadcast is well formed."""
    if keys is None:
        keys = [k for k, v in reference.items() if torch.is_tensor(v)]
    shapes, offsets, kinds = [], [], []
    pos, r = 0, 1
    for k in keys:
        t = reference[k]
        kind = int(num_groups is not None and "lora_S" in k and t.dim() == 2 and t.shape[0] == num_groups)
        if kind:
            if r not in (1, t.shape[1]):
                raise ValueError("all group-weighted lora_S tensors must share one rank")
            r = t.shape[1]
        shapes.append(t.shape)
        offsets.append(pos)
        kinds.append(kind)
        pos += t.numel()
    return FlatSpec(list(keys), shapes, offsets, kinds, pos, num_groups or 1, r)


def pack(spec: FlatSpec, sd: Dict[str, torch.Tensor], device, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    flat = torch.empty(spec.numel, device=device, dtype=torch.float32) if out is None else out
    for k, off, shp in zip(spec.keys, spec.offsets, spec.shapes):
        n = torch.Size(shp).numel()
        flat[off:off + n].copy_(sd[k].detach().reshape(-1), non_blocking=True)      # casts integer tensors to fp32
    return flat


def unpack(spec: FlatSpec, flat: torch.Tensor, like: Optional[Dict[str, torch.Tensor]] = None):
    out = {}
    for k, off, shp in zip(spec.keys, spec.offsets, spec.shapes):
        t = flat[off:off + torch.Size(shp).numel()].view(shp)
        if like is not None:
            # an averaged integer tensor stays a float tensor, as upstream (int64 * python float -> float32)
            t = t.to(device=like[k].device, dtype=like[k].dtype if like[k].is_floating_point() else torch.float32)
        out[k] = t
    return out


def _device_for(tensors) -> torch.device:
    for t in tensors:
        if t.is_cuda:
            return t.device
    if not torch.cuda.is_available():
        raise RuntimeError("fairfedmed_b200.fed_utils needs a CUDA device: aggregation runs on the GPU "
                           "(no CPU fallback; the CPU oracle is oracle/ref_port.average_weights_ema)")
    return torch.device("cuda", torch.cuda.current_device())


def average_weights_EMA(w_g, w, idxs_users, datanumber_client, datanumber_client_by_attr, epoch, max_epoch,
                        beta=0.999, islist=False, shared_half_s=False):
    """Drop-in for utils/fed_utils.py:42-100 (dict branch): every tensor of the state dict is averaged, BatchNorm
    running statistics and batch counters included."""
    if islist:
        raise NotImplementedError("list-of-tensors aggregation is not on the FairLoRA path")
    first = w[idxs_users[0]]
    dev = _device_for(first.values())
    G = None
    by_attr = None
    if datanumber_client_by_attr is not None:
        by_attr = torch.tensor(datanumber_client_by_attr, dtype=torch.float64)
        G = by_attr.shape[1]
        total_by_attr = by_attr[list(idxs_users)].sum(0)
    total = sum(datanumber_client[k] for k in idxs_users)
    spec = build_spec(first, num_groups=G)
    kind, off, ln = spec.segment_tensors(dev)
    acc = None
    for k in idxs_users:
        flat = pack(spec, w[k], dev)
        w_scalar = datanumber_client[k] / total
        if by_attr is not None:
            w_group = (by_attr[k] / total_by_attr).to(torch.float32).to(dev)
        else:
            w_group = torch.ones(1, device=dev)
        scaled = ops.fedavg_scale(flat, kind, off, ln, w_scalar, w_group, spec.G, spec.r)
        acc = scaled if acc is None else acc.add_(scaled)
    beta_decay = beta * (epoch / max(max_epoch, 1))
    prev = pack(spec, w_g, dev)
    new = ops.fedavg_epilogue(acc, prev, kind, off, ln, beta_decay, bool(shared_half_s and by_attr is not None),
                              spec.G, spec.r)
    out = {k: v for k, v in first.items() if k not in spec.keys}
    out.update(unpack(spec, new, like=first))
    return out


def average_weights(w, idxs_users, datanumber_client, datanumber_client_by_attr=None, islist=False):
    """Drop-in for utils/fed_utils.py:6-40: the weighted average without EMA / shared half."""
    zero = {k: torch.zeros_like(v) for k, v in w[idxs_users[0]].items()}
    return average_weights_EMA(zero, w, idxs_users, datanumber_client, datanumber_client_by_attr, 0, 1,
                               islist=islist, shared_half_s=False)


class FederatedAggregator:
    """One simulated site per rank: FedAvg of U, V (scalar weights) and s_g (per-group weights) by all-reduce.

    `aggregate` = exchange_counts (tiny all-reduce fixing the client weights) -> _scale (CUDA) -> all-reduce of the
    flat adapter buffer (NCCL over NVLink / NVSwitch) -> _epilogue (CUDA).  The host logic is backend-agnostic so the
    CPU test-suite can drive it over gloo with a test-side subclass that overrides the two kernel hooks."""

    def __init__(self, spec: FlatSpec, group=None):
        self.spec = spec
        self.group = group

    def exchange_counts(self, n_k: int, n_kg: Optional[Sequence[int]], selected: bool, device):
        """Returns (w_scalar, w_group [G] fp32 on `device`): n_k / sum n_k and n_{k,g} / sum_k n_{k,g} over the
        SELECTED clients (utils/fed_utils.py:58-67); zeros for a client that was not sampled this round."""
        G = self.spec.G
        counts = torch.zeros(1 + G, device=device, dtype=torch.float64)
        if selected:
            counts[0] = float(n_k)
            if n_kg is not None:
                counts[1:] = torch.tensor([float(v) for v in n_kg], dtype=torch.float64)
        if dist.is_initialized():
            dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=self.group)
        totals = counts.cpu()
        w_scalar = (float(n_k) / float(totals[0])) if selected else 0.0
        if selected and n_kg is not None:
            w_group = (torch.tensor([float(v) for v in n_kg], dtype=torch.float64) / totals[1:]).to(torch.float32)
        else:
            w_group = torch.zeros(G, dtype=torch.float32)
        return w_scalar, w_group.to(device)

    def _scale(self, local_flat, w_scalar, w_group):
        kind, off, ln = self.spec.segment_tensors(local_flat.device)
        return ops.fedavg_scale(local_flat, kind, off, ln, w_scalar, w_group, self.spec.G, self.spec.r)

    def _epilogue(self, summed, prev_global_flat, beta_decay, shared_half_s):
        kind, off, ln = self.spec.segment_tensors(summed.device)
        return ops.fedavg_epilogue(summed, prev_global_flat, kind, off, ln, beta_decay, shared_half_s, self.spec.G,
                                   self.spec.r)

    def aggregate(self, local_flat: torch.Tensor, prev_global_flat: torch.Tensor, n_k: int,
                  n_kg: Optional[Sequence[int]], selected: bool, epoch: int, max_epoch: int, beta: float = 0.999,
                  shared_half_s: bool = False) -> torch.Tensor:
        w_scalar, w_group = self.exchange_counts(n_k, n_kg, selected, local_flat.device)
        scaled = self._scale(local_flat, w_scalar, w_group)
        if dist.is_initialized():
            dist.all_reduce(scaled, op=dist.ReduceOp.SUM, group=self.group)
        beta_decay = beta * (epoch / max(max_epoch, 1))
        return self._epilogue(scaled, prev_global_flat, beta_decay, bool(shared_half_s and n_kg is not None))
