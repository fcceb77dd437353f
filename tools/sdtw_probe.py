import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dae.soft_dtw_cuda import softdtw_forward, softdtw_backward
def t(fn, it=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts=[]
    for _ in range(it):
        s,e=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    return sorted(ts)[len(ts)//2]
for (B,N,M) in [(1,32,4096),(1,32,8192),(1,64,4096),(1,128,4096),(1,256,4096),(1,1024,4096),(8,32,4096),(64,32,4096),(592,32,4096),(1184,32,4096)]:
    D=torch.rand(B,N,M,device='cuda')
    f=t(lambda: softdtw_forward(D,1.0,0.0))
    _,W,_=softdtw_forward(D,1.0,0.0); go=torch.ones(B,device='cuda')
    b=t(lambda: softdtw_backward(W,go))
    print(B,N,M,"fwd ms %.4f"%f,"bwd ms %.4f"%b, "fwd cyc/step %.0f"%(f*1e-3*1.965e9/(M+31)), flush=True)
