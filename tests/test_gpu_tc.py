"""Kernel-level checks of the bf16 tcgen05 building blocks (cope_tc_pack / cope_tc_gemm / cope_tc_wgrad) against
fp32 matmuls of the SAME bf16-rounded operands (so only accumulation order differs)."""
import pytest
import torch

from cope_nerf_b200 import _lib as L
from conftest import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"


def pack(W, Np, Kp, transposed=False):
    out_dim, in_dim = W.shape
    n_src, k_src = (in_dim, out_dim) if transposed else (out_dim, in_dim)
    out = torch.empty(Np * Kp, dtype=torch.bfloat16, device=DEV)
    L.call("cope_tc_pack", L.ptr(W), in_dim, n_src, k_src, Np, Kp, int(transposed), L.ptr(out), L.stream())
    return out


@pytest.mark.parametrize("M,N,K,Np,Kp", [(128, 256, 256, 256, 256), (1000, 256, 256, 256, 256), (20000, 204, 256, 208, 256),
                                         (777, 256, 52, 256, 64), (4096, 3, 256, 16, 256), (333, 256, 291, 256, 320),
                                         (131072, 256, 256, 256, 256)])
def test_tc_gemm_store(M, N, K, Np, Kp):
    torch.manual_seed(M + N + K)
    W = torch.randn(N, K, device=DEV) * 0.1
    A = torch.zeros(M, Kp, device=DEV)
    A[:, :K] = torch.randn(M, K, device=DEV)
    Ab = A.to(torch.bfloat16)
    bias = torch.randn(Np, device=DEV)
    Bp = pack(W, Np, Kp)
    out = torch.full((M, Np), 7.0, device=DEV)
    L.call("cope_tc_gemm", M, Np, Kp, L.ptr(Ab), Kp, L.ptr(Bp), L.ptr(bias), 0, 1.0, L.ptr(out), Np, 1, L.stream())
    ref = Ab[:, :K].float() @ W.to(torch.bfloat16).float().t() + bias[:N]
    torch.cuda.synchronize()
    assert rel_err(out[:, :N], ref) < 2e-5
    # bf16 output + softplus epilogue
    outb = torch.empty(M, Np, dtype=torch.bfloat16, device=DEV)
    L.call("cope_tc_gemm", M, Np, Kp, L.ptr(Ab), Kp, L.ptr(Bp), L.ptr(bias), 1, 0.5, L.ptr(outb), Np, 0, L.stream())
    refb = 0.5 * torch.nn.functional.softplus(ref, beta=100)
    assert rel_err(outb[:, :N].float(), refb) < 5e-3


def test_tc_gemm_transposed_pack():
    torch.manual_seed(5)
    M, out_dim, in_dim = 3000, 256, 204
    W = torch.randn(out_dim, in_dim, device=DEV) * 0.1
    A = torch.randn(M, 256, device=DEV).to(torch.bfloat16)
    Bp = pack(W, 208, 256, transposed=True)          # B[n = in][k = out]
    out = torch.empty(M, 208, device=DEV)
    L.call("cope_tc_gemm", M, 208, 256, L.ptr(A), 256, L.ptr(Bp), None, 0, 1.0, L.ptr(out), 208, 1, L.stream())
    ref = A.float() @ W.to(torch.bfloat16).float()
    assert rel_err(out[:, :in_dim], ref) < 2e-5


@pytest.mark.parametrize("P,m,n,Mp,Np", [(64, 256, 256, 256, 256), (5000, 256, 256, 256, 256), (131072, 256, 256, 256, 256),
                                         (9999, 204, 256, 256, 256), (4097, 256, 52, 256, 64), (3000, 3, 256, 128, 256)])
def test_tc_wgrad(P, m, n, Mp, Np):
    torch.manual_seed(P + m + n)
    X = torch.zeros(P, Mp, device=DEV); X[:, :m] = torch.randn(P, m, device=DEV)
    Y = torch.zeros(P, Np, device=DEV); Y[:, :n] = torch.randn(P, n, device=DEV)
    Xb, Yb = X.to(torch.bfloat16), Y.to(torch.bfloat16)
    dW = torch.ones(m, n, device=DEV)
    ws = torch.empty(L.query("cope_tc_wgrad_ws_floats"), device=DEV)
    L.call("cope_tc_wgrad", P, Mp, Np, m, n, L.ptr(Xb), Mp, L.ptr(Yb), Np, L.ptr(dW), n, L.ptr(ws), L.stream())
    ref = Xb[:, :m].float().t() @ Yb[:, :n].float() + 1.0
    assert rel_err(dW, ref) < 2e-5
