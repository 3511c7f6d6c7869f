import sys, torch
sys.path.insert(0, '.')
from cope_nerf_b200 import _lib as L
dev='cuda'
def pack(W,Np,Kp):
    out=torch.empty(Np*Kp,dtype=torch.bfloat16,device=dev)
    L.call("cope_tc_pack",L.ptr(W),W.shape[1],W.shape[0],W.shape[1],Np,Kp,0,L.ptr(out),L.stream()); return out
def timeit(fn,n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n*1000
M=131072
for K in (64,256):
  A=torch.randn(M,K,device=dev).to(torch.bfloat16)
  for N in (16,64,128,256):
    W=torch.randn(N,K,device=dev)*0.1; Bp=pack(W,N,K); bias=torch.zeros(N,device=dev)
    for epi,f32,name in ((0,1,'store_f32'),(0,0,'store_bf16'),(1,0,'softplus_bf16'),(2,0,'relu_bf16')):
        out=torch.empty(M,N,dtype=torch.float32 if f32 else torch.bfloat16,device=dev)
        us=timeit(lambda: L.call("cope_tc_gemm",M,N,K,L.ptr(A),K,L.ptr(Bp),L.ptr(bias),epi,1.0,L.ptr(out),N,f32,L.stream()))
        print(f"K={K} N={N} {name:14s} {us:8.1f} us  {2*M*N*K/us/1e6:8.1f} TFLOP/s")
for Mx in (18944, 18944*2, 18944*4, 18944*7):
    A=torch.randn(Mx,256,device=dev).to(torch.bfloat16); W=torch.randn(16,256,device=dev)*0.1; Bp=pack(W,16,256); bias=torch.zeros(16,device=dev)
    out=torch.empty(Mx,16,dtype=torch.bfloat16,device=dev)
    us=timeit(lambda: L.call("cope_tc_gemm",Mx,16,256,L.ptr(A),256,L.ptr(Bp),L.ptr(bias),2,1.0,L.ptr(out),16,0,L.stream()))
    print(f"Msweep N=16 M={Mx} tiles/CTA={Mx//18944}: {us:.1f} us")
X=torch.randn(M,256,device=dev).to(torch.bfloat16); Y=torch.randn(M,256,device=dev).to(torch.bfloat16); dW=torch.zeros(256,256,device=dev); ws=torch.empty(L.query("cope_tc_wgrad_ws_floats"),device=dev)
us=timeit(lambda: L.call("cope_tc_wgrad",M,256,256,256,256,L.ptr(X),256,L.ptr(Y),256,L.ptr(dW),256,L.ptr(ws),L.stream()))
print(f"wgrad 256x256 P={M}: {us:.1f} us {2*M*256*256/us/1e6:.1f} TFLOP/s")
