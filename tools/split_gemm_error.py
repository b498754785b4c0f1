"""CPU arithmetic check for the planned fp32 mode of the latent projection (DESIGN §6, next target 4): how close is
x.w computed from two-term bf16 splits (x1 w1 + x2 w1 + x1 w2, fp32 accumulation — what ONE K-concatenated bf16
tensor-core GEMM over [X1 | X2 | X1] . [W1 | W1 | W2]^T delivers) to the fp64 product, at the latent layer's shape?
usage: python tools/split_gemm_error.py"""
import torch
torch.manual_seed(0)
N, D, L4 = 2048, 768, 512
X = torch.tanh(torch.randn(N, D))
W = (torch.rand(L4, D) * 2 - 1) / D ** 0.5
ref = X.double() @ W.double().T


def split(t):
    h = t.bfloat16().float()
    return h, (t - h).bfloat16().float()


def rel(a):
    return ((a.double() - ref).norm() / ref.norm()).item()


X1, X2 = split(X)
W1, W2 = split(W)
print(f"bf16 operands (x1 w1)                 rel err {rel(X1 @ W1.T):.2e}")
print(f"two-term split, 3 products            rel err {rel(X1 @ W1.T + X2 @ W1.T + X1 @ W2.T):.2e}")
print(f"two-term split, 4 products            rel err {rel(X1 @ W1.T + X2 @ W1.T + X1 @ W2.T + X2 @ W2.T):.2e}")
print(f"plain fp32 GEMM                       rel err {rel(X @ W.T):.2e}")
