import sys, torch
sys.path.insert(0, '.')
from cope_nerf_b200 import _lib as L
dev='cuda'
M,K=131072,256
A=torch.randn(M,K,device=dev).to(torch.bfloat16)
for N in (16,256):
    W=torch.randn(N,K,device=dev)*0.1
    Bp=torch.empty(N*K,dtype=torch.bfloat16,device=dev)
    L.call("cope_tc_pack",L.ptr(W),K,N,K,N,K,0,L.ptr(Bp),L.stream())
    bias=torch.zeros(N,device=dev); out=torch.empty(M,N,dtype=torch.bfloat16,device=dev)
    for _ in range(3):
        L.call("cope_tc_gemm",M,N,K,L.ptr(A),K,L.ptr(Bp),L.ptr(bias),2,1.0,L.ptr(out),N,0,L.stream())
torch.cuda.synchronize(); print("ok")
