import sys, torch
sys.path.insert(0, '.')
from cope_nerf_b200 import _lib as L
dev='cuda'
M,K=132608,256
A=torch.randn(M,K,device=dev).to(torch.bfloat16)
for N in (16,256):
    W=torch.randn(N,K,device=dev)*0.1
    Bp=torch.empty(N*K,dtype=torch.bfloat16,device=dev)
    L.call("cope_tc_pack",L.ptr(W),K,N,K,N,K,0,L.ptr(Bp),L.stream())
    out=torch.empty(M,N,dtype=torch.bfloat16,device=dev)
    dbg=torch.zeros(6*64,dtype=torch.int64,device=dev)
    for _ in range(3):
        L.call("cope_tc_gemm",M,N,K,L.ptr(A),K,L.ptr(Bp),L.ptr(dbg.view(torch.float32)),2,-1.0,L.ptr(out),N,0,L.stream())
    torch.cuda.synchronize()
    d=dbg.cpu().view(6,64); t0=d[0,0].item()
    names=["prod:got_empty","prod:issued","mma:got_full","mma:committed","epi:got_acc","epi:done"]
    print("N=",N)
    for i,nm in enumerate(names):
        row=[(v-t0) for v in d[i].tolist() if v>0]
        print(f"  {nm:16s}", row[:30])
