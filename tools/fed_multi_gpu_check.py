"""Multi-GPU check (run under torchrun, one rank per GPU): one federated round with one simulated site per rank.
The NCCL-aggregated global adapter buffer must equal the oracle's average_weights_EMA applied to the gathered local
buffers (fp32 tolerance).  Usage:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      tools/fed_multi_gpu_check.py
"""
import os, sys
from pathlib import Path
import torch, torch.distributed as dist
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import fairfedmed_b200.trainer  # noqa: F401,E402
from fairfedmed_b200 import fed_utils  # noqa: E402
from fairfedmed_b200.config import get_cfg_default  # noqa: E402
from fairfedmed_b200.federated import run_federated  # noqa: E402
from fairfedmed_b200.registry import build_trainer  # noqa: E402
from oracle import ref_port as rp  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cfg = get_cfg_default()
cfg.MODEL_ARCH.merge_from_dict(dict(VISION_LAYERS=2, VISION_WIDTH=128, TEXT_LAYERS=2, TEXT_WIDTH=64, TEXT_HEADS=2, EMBED=64))
cfg.INPUT.SIZE = (64, 64)
cfg.DATASET.merge_from_dict(dict(USERS=world, NUM_TRAIN_PER_CLIENT=16 + 8 * 0, NUM_TEST_PER_CLIENT=32))
cfg.DATALOADER.TRAIN_X.BATCH_SIZE = 8
cfg.TRAINER.GLP_OT.OT = "Sinkhorn"
tr = build_trainer(cfg)
tr.step_auc = False
start = tr.get_flat().clone()
dist.broadcast(start, src=0)
# hand-driven round: train locally from the common start, gather, aggregate with the oracle on the CPU
tr.set_flat(start)
tr.train(idx=rank, global_epoch=0, is_fed=True)
local_flat = tr.get_flat().clone()
gathered = [torch.empty_like(local_flat) for _ in range(world)]
dist.all_gather(gathered, local_flat)
spec = tr.flat_spec
n_k = [len(tr.fed_train_loader_x_dict[k].dataset) for k in range(world)]
n_kg = [tr.fed_train_loader_x_dict[k].dataset.count_by_attribute("race") for k in range(world)]
w = [{k: v.cpu() for k, v in fed_utils.unpack(spec, f).items()} for f in gathered]
w_g = {k: v.cpu() for k, v in fed_utils.unpack(spec, start).items()}
ref = rp.average_weights_ema(w_g, w, list(range(world)), n_k, n_kg, 0, 1, shared_half_s=True)
# the product path: same round through run_federated (fresh identical trainer => identical local training)
tr2 = build_trainer(cfg)
tr2.step_auc = False
_, global_flat, hist = run_federated(cfg, rounds=1, shared_half_s=True, trainer=tr2, log=lambda *a: None)
got = fed_utils.unpack(spec, global_flat)
worst = 0.0
for k in spec.keys:
    err = float((got[k].cpu() - ref[k]).abs().max() / (ref[k].abs().max() + 1e-12))
    worst = max(worst, err)
# every rank must hold bit-identical global weights after the all-reduce
chk = [torch.empty_like(global_flat) for _ in range(world)]
dist.all_gather(chk, global_flat)
same = all(torch.equal(chk[0], c) for c in chk)
if rank == 0:
    print(f"world={world} aggregated-vs-oracle worst rel err {worst:.3e}; identical across ranks: {same}; "
          f"round auc {hist[0]['auc']:.2f}")
    assert worst < 1e-3 and same
    print("MULTI-GPU FEDERATED CHECK OK")
dist.destroy_process_group()
