"""One fwd + bwd of torch SDPA's cuDNN backend on the bench workload (config 3), for an ncu launch list: which kernels the
sm_100-native competitor runs (count, grid, cluster, shared memory, registers, tensor activity)."""
import torch
from torch.nn.attention import SDPBackend, sdpa_kernel
B, H, N, D = 2, 32, 8192, 128
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(42)
q, k, v, do = (torch.randn(B, H, N, D, generator=g).to(torch.bfloat16).to(dev) for _ in range(4))
q.requires_grad_(True), k.requires_grad_(True), v.requires_grad_(True)
for _ in range(3):
    with sdpa_kernel(SDPBackend.CUDNN_ATTENTION):
        o = torch.nn.functional.scaled_dot_product_attention(q, k, v, scale=D ** -0.5, is_causal=True)
    o.backward(do)
    q.grad = k.grad = v.grad = None
torch.cuda.synchronize()
print("ok")
