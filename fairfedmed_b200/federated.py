"""Federated round loop — the `FedOTPLoRA` branch of federated_main.main (federated_main.py:604-690).

Two layouts, same arithmetic:
  * sequential (world size 1): all K simulated clients run one after the other on one GPU, like the reference;
  * sharded (world size K): one client per rank / B200, the per-round aggregation is an NCCL all-reduce of the
    flat adapter buffer (fed_utils.FederatedAggregator).  Clients are independent inside a round, so this is
    weak scaling in the number of sites; optimizer state is per client here (the reference leaks one shared
    momentum buffer across its sequential clients — SURVEY.md F7; round-0 aggregated weights with zero initial
    momentum are identical in both layouts).
Every client starts a round from the global weights (federated_main.py:645-652 with the default empty
`idxs_users_train`).
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch
import torch.distributed as dist

from . import ops
from .fed_utils import FederatedAggregator
from .registry import build_trainer


def _client_counts(trainer, k: int):
    ds = trainer.fed_train_loader_x_dict[k].dataset
    n_kg = None
    if not trainer.cfg.TRAINER.GLP_OT_LORA.DISABLE_ATTR:
        n_kg = ds.count_by_attribute(trainer.cfg.DATASET.ATTRIBUTE_TYPE)
    return len(ds), n_kg


def select_clients(epoch: int, n_users: int, frac: float, rng: np.random.RandomState) -> List[int]:
    """All users in round 0, then max(int(frac * K), 1) sampled without replacement (federated_main.py:606-613)."""
    if epoch == 0:
        return list(range(n_users))
    m = max(int(frac * n_users), 1)
    return sorted(rng.choice(range(n_users), m, replace=False).tolist())


def run_federated(cfg, rounds: Optional[int] = None, frac: float = 1.0, shared_half_s: bool = True,
                  evaluate: bool = True, trainer=None, log=print, beta: float = 0.999,
                  max_epoch: Optional[int] = None, keep_locals: bool = False):
    """Returns (trainer, global_flat, history). Sharded when torch.distributed is initialised with world > 1.

    `rounds` = how many rounds to run now; `max_epoch` = the EMA horizon of average_weights_EMA (cfg.OPTIM.ROUND in the
    reference, federated_main.py:631-633) — they differ when a run is cut short; `beta` as utils/fed_utils.py:42.
    `keep_locals`: history entries also carry this process's per-client buffers of the round (parity checks)."""
    rounds = cfg.OPTIM.ROUND if rounds is None else rounds
    max_epoch = rounds if max_epoch is None else max_epoch
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    K = cfg.DATASET.USERS
    if world > 1 and world != K:
        raise ValueError(f"sharded layout needs one rank per client: world={world}, DATASET.USERS={K}")
    trainer = trainer or build_trainer(cfg)
    trainer.fed_before_train()
    rng = np.random.RandomState(cfg.SEED)          # identical on every rank: same client sample everywhere
    spec = trainer.flat_spec
    dev = trainer.device
    kind, off, ln = spec.segment_tensors(dev)
    counts = [_client_counts(trainer, k) for k in range(K)]
    global_flat = trainer.get_flat().clone()
    if world > 1:
        dist.broadcast(global_flat, src=0)          # every site starts from rank 0's initialisation
    agg = FederatedAggregator(spec)
    history = []
    # ONE StepLR is shared by all clients upstream and stepped by every client-epoch in sequence (F6/F7): client number
    # j of a round trains at the rate reached after all earlier client-epochs.  Both layouts reproduce that count.
    per_client = trainer._n_opt_steps() * trainer.max_epoch
    sched_base = trainer.sched_steps
    for epoch in range(rounds):
        idxs = select_clients(epoch, K, frac, rng)
        locals_kept = {}
        if world == 1:
            total = sum(counts[k][0] for k in idxs)
            tot_g = None
            if counts[0][1] is not None:
                tot_g = torch.tensor([counts[k][1] for k in idxs], dtype=torch.float64).sum(0)
            acc = None
            for j, k in enumerate(idxs):
                trainer.set_flat(global_flat)
                trainer.sched_steps = sched_base + j * per_client
                trainer.lr_dev.fill_(float(trainer.current_lr()))
                trainer.train(idx=k, global_epoch=epoch, is_fed=True, is_last_client=(k == idxs[-1]))
                w_group = (torch.tensor(counts[k][1], dtype=torch.float64) / tot_g).float().to(dev) \
                    if tot_g is not None else torch.zeros(spec.G, device=dev)
                if keep_locals:
                    locals_kept[k] = trainer.get_flat().clone()
                scaled = ops.fedavg_scale(trainer.get_flat(), kind, off, ln, counts[k][0] / total, w_group, spec.G,
                                          spec.r)
                acc = scaled if acc is None else acc.add_(scaled)
            beta_decay = beta * (epoch / max(max_epoch, 1))
            global_flat = ops.fedavg_epilogue(acc, global_flat, kind, off, ln, beta_decay,
                                              bool(shared_half_s and tot_g is not None), spec.G, spec.r)
        else:
            selected = rank in idxs
            trainer.set_flat(global_flat)
            if selected:
                trainer.sched_steps = sched_base + idxs.index(rank) * per_client
                trainer.lr_dev.fill_(float(trainer.current_lr()))
                trainer.train(idx=rank, global_epoch=epoch, is_fed=True, is_last_client=(rank == idxs[-1]))
                if keep_locals:
                    locals_kept[rank] = trainer.get_flat().clone()
            global_flat = agg.aggregate(trainer.get_flat(), global_flat, counts[rank][0], counts[rank][1], selected,
                                        epoch, max_epoch, beta=beta, shared_half_s=shared_half_s)
        sched_base += len(idxs) * per_client
        trainer.sched_steps = sched_base
        trainer.lr_dev.fill_(float(trainer.current_lr()))
        entry = {"round": epoch, "clients": idxs}
        if keep_locals:
            entry["locals"] = locals_kept
        if evaluate:
            trainer.set_flat(global_flat)
            mine = range(K) if world == 1 else [rank]
            res = [trainer.test(idx=k, current_epoch=epoch) for k in mine]
            # "Global test acc / error / macro_f1 / auc" = mean over ALL clients (federated_main.py:677-690)
            sums = torch.tensor([[r[i] for i in range(4)] for r in res], dtype=torch.float64, device=dev).sum(0)
            if world > 1:
                dist.all_reduce(sums, op=dist.ReduceOp.SUM)
            sums = (sums / K).tolist()
            entry["acc"], entry["err"], entry["f1"], entry["auc"] = (float(v) for v in sums)
            if rank == 0:
                log(f"[round {epoch}] clients {idxs} acc {entry['acc']:.2f} auc {entry['auc']:.2f}")
        history.append(entry)
    trainer.set_flat(global_flat)
    trainer.fed_after_train()
    return trainer, global_flat, history
