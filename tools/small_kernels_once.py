"""Launch each of the small memory-bound kernels twice at its bench size (for an `ncu --set full` capture that stays small).
  ncu --set full --clock-control none -k regex:'lora_merge|oct_|patchify|vit_embed|sim_kernel|head_bwd' -o out python tools/small_kernels_once.py
"""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from fairfedmed_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
mean3 = torch.tensor([0.48145466, 0.4578275, 0.40821073], device=dev)
std3 = torch.tensor([0.26862954, 0.26130258, 0.27577711], device=dev)
Wm = torch.randn(2048, 2048, generator=g).to(dev)
Am = (0.1 * torch.randn(2048, 32, generator=g)).to(dev)
Bmm = torch.randn(32, 2048, generator=g).to(dev)
yo = (3.0 * torch.randn(256, 3, 224, 224, generator=g)).to(dev)
img = torch.randint(0, 256, (64, 1, 224, 224), generator=g).float().repeat(1, 3, 1, 1).to(dev)
pe = torch.randn(64, 196, 768, generator=g).bfloat16().to(dev)
cls_e = torch.randn(768, generator=g).to(dev)
pos_e = torch.randn(197, 768, generator=g).to(dev)
one, zero = torch.ones(768, device=dev), torch.zeros(768, device=dev)
feats = torch.randn(64, 197, 512, generator=g).bfloat16().to(dev).requires_grad_(True)
txt = torch.randn(4, 512, generator=g).to(dev).requires_grad_(True)
ls = torch.tensor(4.6, device=dev, requires_grad=True)
dl = torch.randn(64, 2, generator=g).to(dev)
for _ in range(2):
    ops.lora_merged_weight_op(Wm, Am, Bmm, 0.25)
    ops.lora_merged_weight_bwd_op(Wm, Am, Bmm, 0.25)
    pt, lo_, hi_ = ops.oct_minmax_patchify_op(yo, mean3, std3, 16)
    ops.oct_input_bwd_op(pt, yo, lo_, hi_, std3, 16)
    ops.patchify_normalize(img, mean3, std3, 16, True)
    ops.vit_embed_ln(pe, cls_e, pos_e, one, zero, one, zero, 1e-5, 1e-5)
    lg, _, _ = ops.ot_head(feats, txt, ls, n_cls=2, num_slices=1, ot="Sinkhorn", batch_first=True)
    (lg * dl).sum().backward()
torch.cuda.synchronize()
print("ok")
