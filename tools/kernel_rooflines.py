"""Isolated device times of every ffm:: kernel at config-2 sizes (and stress sizes for the tiny ones), against the
roofline that bounds it.  CUPTI kernel durations via torch.profiler, L2 flushed (256 MB memset) between repetitions so
HBM-bound kernels with < 126 MB working sets are not served from L2.  Run on the B200 box:
    python tools/kernel_rooflines.py > gpurun_out/kernel_rooflines.json
"""
import collections, json, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from fairfedmed_b200 import ops  # noqa: E402
from torch.profiler import profile, ProfilerActivity  # noqa: E402

dev = torch.device("cuda:0")
peaks = json.loads((Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json").read_text()) \
    if (Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json").exists() else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
HBM, TF = peaks["hbm_gbs"], peaks["bf16_tflops"]
flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
g = torch.Generator().manual_seed(0)


def measure(fn, reps=10, flush=True):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(reps):
            if flush:
                flush_buf.zero_()
                flush_buf.view(torch.int32).sum()      # read pass: the timed kernel does not pay for dirty write-backs
            fn()
        torch.cuda.synchronize()
    agg = collections.defaultdict(list)
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA and "ffm::" in ev.name:
            agg[ev.name.split("(")[0].replace("void ", "")].append(ev.device_time)
    return {k: sum(v) / len(v) for k, v in agg.items()}     # us per launch


rows = []


def add(kernel, what, us, bytes_=None, flops=None, note=""):
    r = {"kernel": kernel, "shape": what, "us": round(us, 2)}
    if bytes_ is not None:
        r.update(bound="hbm", algorithmic_MB=round(bytes_ / 1e6, 2), achieved_GBs=round(bytes_ / us / 1e3, 1),
                 frac_of_measured_hbm=round(bytes_ / us / 1e3 / HBM, 3))
    if flops is not None:
        r.update(bound="tensor", algorithmic_GFLOP=round(flops / 1e9, 2), achieved_TFLOPs=round(flops / us / 1e6, 1),
                 frac_of_measured_bf16_burst=round(flops / us / 1e6 / TF, 3))
    if note:
        r["note"] = note
    rows.append(r)


# ---------------- fused SVLoRA GEMM + adapter-gradient kernels (config 2: T = 197*64) ----------------
T, r, B = 197 * 64, 12, 64
for (K, N, act) in [(768, 3072, 1), (3072, 768, 0)]:
    x = torch.randn(T, K, generator=g).bfloat16().to(dev)
    W = (torch.randn(N, K, generator=g) * K ** -0.5).bfloat16().to(dev)
    Wt = W.t().contiguous()
    bias = torch.randn(N, generator=g).to(dev)
    A = (0.05 * torch.randn(K, r, generator=g)).to(dev)
    Bm = torch.randn(r, N, generator=g).to(dev)
    s_eff = (torch.rand(B, r, generator=g) + 0.1).to(dev)
    dy = (0.1 * torch.randn(T, N, generator=g)).bfloat16().to(dev)
    out = {}

    def fwd():
        out["f"] = ops.svlora_fwd(x, W, bias, A, Bm, s_eff, 1 / 6, B, 1, act, 197)
    m = measure(fwd)
    fl = 2.0 * T * K * N + 2.0 * T * r * (K + N)
    for k, us in m.items():
        if "gemm" in k:
            add(k, f"fwd T={T} K={K} N={N} act={act}", us, flops=fl)
        elif "prep" in k:
            add(k, f"K={K} N={N}", us, note="latency bound (adapter tiles for both directions)")
    y, y_dact, h, z, tiles = out["f"]
    aux = torch.rand(T, K, generator=g).bfloat16().to(dev) if not act else None   # dX of c_proj multiplies by QuickGELU'

    def bwd():
        ops.svlora_bwd(dy, x, Wt, A, Bm, s_eff, h, z, tiles, aux, 1 / 6, B, 1, 197)
    m = measure(bwd)
    for k, us in m.items():
        if "gemm" in k:
            add(k, f"dX T={T} K={N} N={K} gelu'={aux is not None}", us, flops=fl)
        elif "adapter_grad_kernel" in k or "adapter_grad_tma_kernel" in k:
            add(k, f"dA+dB T={T} K={K} N={N}", us, bytes_=T * (K + N) * 2.0 + 2 * T * 16 * 2.0)
        elif "finalize" in k:
            add(k, f"K={K} N={N} nS={B}", us, note="latency bound (partials fold + segmented ds_eff)")

# ---------------- frozen in_proj / out_proj of the image tower (adapter-free build of the fused GEMM) ----------------
for (K, N) in [(768, 2304), (768, 768), (2304, 768)]:
    x = torch.randn(T, K, generator=g).bfloat16().to(dev)
    W = (torch.randn(N, K, generator=g) * K ** -0.5).bfloat16().to(dev)
    bias = torch.randn(N, generator=g).to(dev)
    m = measure(lambda: ops.frozen_linear_op(x, W, bias))
    for k, us in m.items():
        add(k, f"frozen linear T={T} K={K} N={N}", us, flops=2.0 * T * K * N)

# ---------------- attention core (tcgen05): image tower B'=64, L=197, 12 heads x 64 ----------------
qkv = torch.randn(64, 197, 2304, generator=g).bfloat16().to(dev)
do = (0.1 * torch.randn(64, 197, 768, generator=g)).bfloat16().to(dev)
st = {}


def att_f():
    st["o"] = ops.attention_fwd(qkv, 12, False, True)
m = measure(att_f)
att_fl = 4.0 * 64 * 12 * 197 * 197 * 64
for k, us in m.items():
    add(k, "fwd B'=64 L=197 H=12 d=64", us, flops=att_fl,
        note=f"HBM floor (qkv + out = {64 * 197 * 3072 * 2 / 1e6:.0f} MB) {64 * 197 * 3072 * 2 / HBM / 1e3:.1f} us")
o_, lse_ = st["o"]
m = measure(lambda: ops.attention_bwd(qkv, o_, do, lse_, 12, False, True))
for k, us in m.items():
    add(k, "bwd B'=64 L=197 H=12 d=64", us, flops=2.5 * att_fl,
        note=f"HBM floor (qkv + out + d_out + d_qkv = {64 * 197 * 6144 * 2 / 1e6:.0f} MB) {64 * 197 * 6144 * 2 / HBM / 1e3:.1f} us")

# ---------------- OCT input side (config 3: 256 slice-images) and merged LoRA weight (RN50 attention pool) ----------------
yo = (3.0 * torch.randn(256, 3, 224, 224, generator=g)).to(dev)
mean3 = torch.tensor([0.48145466, 0.4578275, 0.40821073], device=dev)
std3 = torch.tensor([0.26862954, 0.26130258, 0.27577711], device=dev)
st = {}


def oct_f():
    st["o"] = ops.oct_minmax_patchify_op(yo, mean3, std3, 16)
m = measure(oct_f)
npx = 256 * 3 * 224 * 224
for k, us in m.items():
    add(k, "B'=256 3x224x224 fp32", us, bytes_=npx * (4.0 if "minmax" in k else 6.0))
pt, lo_, hi_ = st["o"]
dpt = torch.randn(pt.shape, generator=g).bfloat16().to(dev)
m = measure(lambda: ops.oct_input_bwd_op(dpt, yo, lo_, hi_, std3, 16))
for k, us in m.items():
    add(k, "B'=256 3x224x224", us, bytes_=npx * 10.0, note="algorithmic bytes = d_patches + y in, d_y out (second pass served by L2)")
Wm = torch.randn(2048, 2048, generator=g).to(dev)
Am = (0.1 * torch.randn(2048, 32, generator=g)).to(dev)
Bmm = torch.randn(32, 2048, generator=g).to(dev)
m = measure(lambda: ops.lora_merged_weight_op(Wm, Am, Bmm, 0.25))
for k, us in m.items():
    add(k, "out=2048 in=2048 r=32 fp32", us, bytes_=2.0 * 2048 * 2048 * 4)
m = measure(lambda: ops.lora_merged_weight_bwd_op(Wm, Am, Bmm, 0.25))
for k, us in m.items():
    add(k, "out=2048 in=2048 r=32 fp32", us, bytes_=2048 * 2048 * 4.0)

# ---------------- residual add + LayerNorm ----------------
xa = torch.randn(T, 768, generator=g).bfloat16().to(dev)
xb = torch.randn(T, 768, generator=g).bfloat16().to(dev)
gam, bet = torch.ones(768, device=dev), torch.zeros(768, device=dev)
st = {}


def ln_f():
    st["o"] = ops.add_layernorm_fwd(xa, xb, gam, bet, 1e-5)
m = measure(ln_f)
for k, us in m.items():
    add(k, f"rows={T} C=768 (+res)", us, bytes_=4.0 * T * 768 * 2)
s_, ln_, mean_, rstd_ = st["o"]
m = measure(lambda: ops.add_layernorm_bwd(xa, xb, s_, gam, mean_, rstd_))
for k, us in m.items():
    add(k, f"rows={T} C=768 (+d_res)", us, bytes_=4.0 * T * 768 * 2)

# ---------------- GLP_OT head + stand-alone Sinkhorn ----------------
feats = torch.randn(197, 64, 512, generator=g).bfloat16().to(dev)
txt = torch.randn(4, 512, generator=g).to(dev)
ls = torch.tensor(4.6, device=dev)
m = measure(lambda: ops.ot_head(feats, txt, ls, n_cls=2, num_slices=1, ot="Sinkhorn"))
for k, us in m.items():
    if "sim_kernel" in k:
        add(k, "M+1=197 B'=64 D=512 bf16", us, bytes_=197 * 64 * 512 * 2.0)
    else:
        add(k, "config 2 head (128 problems)", us, note="latency bound: 200 KB problem")
# head forward + backward in the tower's own batch-first layout
feats_bf = feats.transpose(0, 1).contiguous().requires_grad_(True)
txt_g = txt.clone().requires_grad_(True)
ls_g = ls.clone().requires_grad_(True)
dl = torch.randn(64, 2, generator=g).to(dev)


def head_fb():
    lg, _, _ = ops.ot_head(feats_bf, txt_g, ls_g, n_cls=2, num_slices=1, ot="Sinkhorn", batch_first=True)
    (lg * dl).sum().backward()
m = measure(head_fb)
for k, us in m.items():
    if "head_bwd" in k:
        add(k, "B'=64 M+1=197 D=512 bf16 (batch-first)", us, bytes_=2 * 197 * 64 * 512 * 2.0,
            note="algorithmic bytes = read features + write their gradient")
    elif "txt_bwd" in k:
        add(k, "296 partials x 4 x 512", us, note="latency bound")

# ---------------- ViT input side ----------------
img = torch.randint(0, 256, (64, 1, 224, 224), generator=g).float().repeat(1, 3, 1, 1).to(dev)
mean = torch.tensor([0.48145466, 0.4578275, 0.40821073], device=dev)
std = torch.tensor([0.26862954, 0.26130258, 0.27577711], device=dev)
m = measure(lambda: ops.patchify_normalize(img, mean, std, 16, True))
for k, us in m.items():
    add(k, "B'=64 3x224x224 fp32 -> bf16 patches", us, bytes_=64 * 3 * 224 * 224 * 6.0)
pe = torch.randn(64, 196, 768, generator=g).bfloat16().to(dev)
cls_e = torch.randn(768, generator=g).to(dev)
pos_e = torch.randn(197, 768, generator=g).to(dev)
one, zero = torch.ones(768, device=dev), torch.zeros(768, device=dev)
m = measure(lambda: ops.vit_embed_ln(pe, cls_e, pos_e, one, zero, one, zero, 1e-5, 1e-5))
for k, us in m.items():
    add(k, "B'=64 G=196 C=768", us, bytes_=(64 * 196 + 2 * 64 * 197) * 768 * 2.0 + 197 * 768 * 4.0)

for P in (128, 8192, 65536):
    sim = torch.rand(P, 196, 2, generator=g)
    Kmat = torch.exp(-(1 - sim) / 0.1).to(dev)
    st = {}

    def sk():
        st["o"] = ops.sinkhorn(Kmat, mode="Sinkhorn")
    m = measure(sk, reps=5)
    iters = int(st["o"][1][0].item())
    for k, us in m.items():
        add(k, f"P={P} M=196 N=2, {iters} iterations", us, bytes_=2.0 * P * 196 * 2 * 4,
            note="algorithmic bytes = read K once + write T once (SURVEY 8d); problems beyond the resident warps are "
                 "re-streamed every iteration because the reference's stopping rule is batch-global")

# ---------------- aggregation / SGD / metrics ----------------
from fairfedmed_b200 import _cabi  # noqa: E402
Pn = 1110880
flat = torch.randn(Pn, generator=g).to(dev)
mom = torch.zeros_like(flat)
grad = torch.randn(Pn, generator=g).to(dev)
m = measure(lambda: ops.sgd_step_(flat, grad, mom, 1e-3, 0.9, 5e-4, 2, False))
for k, us in m.items():
    add(k, f"P={Pn} fp32, double step", us, bytes_=5.0 * Pn * 4)
# FedAvg over the flat buffer: scale (per-client weights) and EMA epilogue
from fairfedmed_b200.fed_utils import FlatSpec  # noqa: E402,F401
kinds = torch.zeros(1, dtype=torch.int32, device=dev)
offs = torch.zeros(1, dtype=torch.int64, device=dev)
lens = torch.full((1,), Pn, dtype=torch.int64, device=dev)
wg = torch.ones(3, device=dev) / 3
try:
    m = measure(lambda: ops.fedavg_scale(flat, kinds, offs, lens, 0.5, wg, 3, 12))
    for k, us in m.items():
        add(k, f"P={Pn} fp32", us, bytes_=2.0 * Pn * 4)
    m = measure(lambda: ops.fedavg_epilogue(flat, grad, kinds, offs, lens, 0.06, True, 3, 12))
    for k, us in m.items():
        add(k, f"P={Pn} fp32", us, bytes_=3.0 * Pn * 4)
except Exception as e:      # argument layout of the aggregation entry points changed: keep the rest of the table
    rows.append({"kernel": "fedavg", "note": f"not measured: {type(e).__name__}: {e}"[:200]})
N = 200000
prob = torch.softmax(torch.randn(N, 2, generator=g), 1).to(dev)
lab = torch.randint(0, 2, (N,), generator=g).to(dev)
att = torch.randint(0, 3, (3, N), generator=g).to(dev)
m = measure(lambda: ops.group_auc_counts(prob, lab, att, 3))
tot = sum(m.values())
add("group_auc (all kernels)", f"N={N}, 3 attributes x 3 groups", tot, bytes_=N * (8 + 4 + 3 * 4.0),
    note="sort-dominated: " + ", ".join(f"{k.split('::')[-1]} {v:.1f}us" for k, v in sorted(m.items(), key=lambda kv: -kv[1])[:4]))

print(json.dumps({"peaks": {"hbm_gbs": HBM, "bf16_tflops_burst": TF}, "rows": rows}, indent=1))
