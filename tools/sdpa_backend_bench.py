import torch, torch.nn.functional as F
from torch.nn.attention import sdpa_kernel, SDPBackend
dev="cuda"
B,L,H,hd=64,197,12,64
qkv=torch.randn(B,L,3,H,hd,device=dev,dtype=torch.bfloat16,requires_grad=True)
def run(backend):
    q,k,v=qkv.unbind(2)
    with sdpa_kernel([backend]):
        out=F.scaled_dot_product_attention(q.transpose(1,2),k.transpose(1,2),v.transpose(1,2))
    return out
for be in (SDPBackend.CUDNN_ATTENTION, SDPBackend.FLASH_ATTENTION, SDPBackend.EFFICIENT_ATTENTION):
    try:
        g=torch.randn(B,H,L,hd,device=dev,dtype=torch.bfloat16)
        for _ in range(3):
            o=run(be); o.backward(g); qkv.grad=None
        torch.cuda.synchronize()
        e0,e1,e2=[torch.cuda.Event(enable_timing=True) for _ in range(3)]
        tf=tb=0
        for _ in range(20):
            e0.record(); o=run(be); e1.record(); o.backward(g); e2.record(); torch.cuda.synchronize()
            tf+=e0.elapsed_time(e1); tb+=e1.elapsed_time(e2); qkv.grad=None
        print(be, f"fwd {tf/20*1e3:.1f} us  bwd(incl cat of dq/dk/dv) {tb/20*1e3:.1f} us")
    except Exception as ex:
        print(be, "failed:", str(ex)[:100])
