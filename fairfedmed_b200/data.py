"""Synthetic FairFedMed-shaped data: no datasets are available offline.

Samples follow the reference's tensor contract (utils/data_utils.py:559-782 through DatasetWrapperAttr,
Dassl/dassl/data/data_manager.py:402-515): a batch is a dict
    {"img": float32 [B, 3|32, 224, 224] with raw 0..255 values, "label": int64 [B], "attrs": int64 [B, n_attr]}.
2-D SLO / chest images are one uint8-valued channel replicated to 3; OCT volumes are [32, 224, 224].
Labels are balanced inside every batch (the reference's per-step AUC needs both classes).
Batches are staged in PINNED host memory so the trainer's H2D copies are asynchronous.
"""
from __future__ import annotations

from typing import List, Sequence

import torch

from .config import ATTRIBUTE_GROUPS


class SyntheticClientDataset:
    def __init__(self, n: int, dataset: str, modality: str, attributes: Sequence[str], resolution: int = 224,
                 seed: int = 1, materialize: bool = True):
        g = torch.Generator().manual_seed(seed)
        self.n = n
        self.attributes = list(attributes)
        self.groups = [len(ATTRIBUTE_GROUPS[dataset][a]) for a in self.attributes]
        self.is_3d = modality in {"oct_bscans", "oct_bscans_3d", "mac_onh", "onh_mac"}
        self.channels = 32 if self.is_3d else 3
        self.resolution = resolution
        self.label = (torch.arange(n) % 2).to(torch.int64)
        self.attrs = torch.stack([torch.randint(0, G, (n,), generator=g) for G in self.groups], dim=1)
        self._seed = seed
        self._gray = None
        if materialize:
            c = self.channels if self.is_3d else 1
            self._gray = torch.randint(0, 256, (n, c, resolution, resolution), generator=g, dtype=torch.uint8)

    def __len__(self):
        return self.n

    def count_by_attribute(self, attr_name: str) -> List[int]:
        """Samples per group of one attribute (DatasetWrapperAttr.count_by_attribute, data_manager.py:435-460)."""
        a = self.attributes.index(attr_name)
        return torch.bincount(self.attrs[:, a], minlength=self.groups[a]).tolist()

    def images(self, idx: torch.Tensor) -> torch.Tensor:
        g = self._gray[idx].to(torch.float32)
        return g if self.is_3d else g.repeat(1, 3, 1, 1)


class SyntheticLoader:
    """Iterates fixed-size batches (drop_last) from a SyntheticClientDataset, pinned when CUDA is present."""

    def __init__(self, dataset: SyntheticClientDataset, batch_size: int, shuffle: bool, seed: int = 0,
                 drop_last: bool = True):
        self.dataset, self.batch_size, self.shuffle, self.drop_last = dataset, batch_size, shuffle, drop_last
        self._g = torch.Generator().manual_seed(seed)
        self._pin = torch.cuda.is_available()

    def __len__(self):
        n = len(self.dataset)
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        n = len(self.dataset)
        if self.shuffle:
            # keep every batch class-balanced: shuffle the two classes separately and interleave
            pos = torch.nonzero(self.dataset.label == 1).flatten()
            neg = torch.nonzero(self.dataset.label == 0).flatten()
            pos = pos[torch.randperm(pos.numel(), generator=self._g)]
            neg = neg[torch.randperm(neg.numel(), generator=self._g)]
            m = min(pos.numel(), neg.numel())
            order = torch.stack([neg[:m], pos[:m]], dim=1).flatten()
        else:
            order = torch.arange(n)
        for i in range(len(self)):
            idx = order[i * self.batch_size:(i + 1) * self.batch_size]
            if idx.numel() == 0:
                break
            batch = {"img": self.dataset.images(idx), "label": self.dataset.label[idx].clone(),
                     "attrs": self.dataset.attrs[idx].clone()}
            if self._pin:
                batch = {k: v.pin_memory() for k, v in batch.items()}
            yield batch


class CachedLoader:
    """Replays a fixed list of (pinned-host) batches: the loader a throughput run uses so that batch SYNTHESIS (uint8 ->
    fp32, channel replication on the host) is not what gets measured; the H2D copy of every batch still happens per
    step in the trainer.  `dataset` is forwarded for len() / count_by_attribute()."""

    def __init__(self, batches, n_batches: int, dataset=None):
        self.batches, self.n, self.dataset = list(batches), int(n_batches), dataset

    def __len__(self):
        return self.n

    def __iter__(self):
        for i in range(self.n):
            yield self.batches[i % len(self.batches)]


class _DatasetInfo:
    classnames = ["NOT Glaucoma", "Glaucoma"]   # an explicit LIST (upstream iterates a set: hash-order hazard)
    lab2cname = {0: "NOT Glaucoma", 1: "Glaucoma"}
    num_classes = 2


class SyntheticDataManager:
    """Stand-in for Dassl's DataManager (data_manager.py:62-201): per-client train / test loaders."""

    def __init__(self, cfg):
        ds = cfg.DATASET
        self.dataset = _DatasetInfo()
        if ds.NAME == "FedChexMimic":
            self.dataset.classnames = ["No Finding", "Finding"]
        self.fed_train_loader_x_dict, self.fed_test_loader_x_dict = {}, {}
        res = cfg.INPUT.SIZE[0]
        for k in range(ds.USERS):
            tr = SyntheticClientDataset(ds.NUM_TRAIN_PER_CLIENT, ds.NAME, ds.MODALITY_TYPE, ds.ATTRIBUTES, res,
                                        seed=cfg.SEED * 1000 + 2 * k)
            te = SyntheticClientDataset(ds.NUM_TEST_PER_CLIENT, ds.NAME, ds.MODALITY_TYPE, ds.ATTRIBUTES, res,
                                        seed=cfg.SEED * 1000 + 2 * k + 1)
            self.fed_train_loader_x_dict[k] = SyntheticLoader(tr, cfg.DATALOADER.TRAIN_X.BATCH_SIZE, True,
                                                              seed=cfg.SEED + k)
            self.fed_test_loader_x_dict[k] = SyntheticLoader(te, cfg.DATALOADER.TEST.BATCH_SIZE, False,
                                                             drop_last=False)
