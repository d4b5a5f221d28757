import sys, torch
sys.path.insert(0, '/root/repo')
from videopainter_b200 import ops
BF16=torch.bfloat16; dev='cuda'
B,H,S,St,D=2,48,17776,226,3072; M=B*S
g=torch.Generator(device=dev).manual_seed(0)
rn=lambda *s, sc=1.0: (torch.randn(*s, device=dev, generator=g)*sc).to(BF16)
x=rn(M,D); w=rn(3*D,D,sc=0.02); b=rn(3*D); nq=(rn(64),rn(64))
q,k,v=(torch.empty(B,H,S,64,dtype=BF16,device=dev) for _ in range(3))
c32=torch.rand(S-St,32,device=dev); s32=torch.rand(S-St,32,device=dev)
cos=c32.repeat_interleave(2,dim=1).contiguous(); sin=s32.repeat_interleave(2,dim=1).contiguous()
pairs=torch.stack([c32,s32],dim=-1).reshape(S-St,64).contiguous()
flush=torch.empty(256<<20,dtype=torch.uint8,device=dev)
def t(fn,n=5):
    for _ in range(3): fn()
    ts=[]
    for _ in range(n):
        flush.zero_(); a,e=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(e))
    return sorted(ts)[n//2]
print('qkv rope', t(lambda: ops.gemm_qkv(x,w,b,M,D,S,H,0,q,k,v,nq,nq,1e-6,(cos,sin),St)))
print('qkv rope compact', t(lambda: ops.gemm_qkv(x,w,b,M,D,S,H,0,q,k,v,nq,nq,1e-6,(cos,sin,pairs),St)))
print('qkv norope', t(lambda: ops.gemm_qkv(x,w,b,M,D,S,H,0,q,k,v,nq,nq,1e-6,None,St)))
o=torch.empty(M,3*D,dtype=BF16,device=dev)
print('plain bias N=9216', t(lambda: ops.gemm_bias(x,w,b,o,M,3*D,D,M,0,0)))
