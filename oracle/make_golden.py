"""Generates tests/golden/*.npz — inputs and ground-truth outputs of the reference's own correctness checks.

The reference holds no golden files; its tests compare against torch computations on seeded inputs.  Those
scripts (src/test_correctness.py, src/flash_attention_openai_tutorial.py::test_op) need a CUDA GPU and Triton at
import time, so the ground-truth halves are restated here line for line and run on the CPU:

  sdpa_*      test_correctness.py:28-33,46-48 — torch.manual_seed(seed); Q,K,V,dO = randn(B,H,N,d) fp32;
              O = scaled_dot_product_attention(Q,K,V,scale=1); dQ,dK,dV = autograd.grad(O,(Q,K,V),dO)
  tutorial_*  flash_attention_openai_tutorial.py:523-549 — torch.manual_seed(20); q,k,v ~ normal(0, 0.5) fp16;
              sm_scale = 0.5; p = q @ k^T * sm_scale; causal tril mask; p = softmax(p.float()).half();
              ref_out = p @ v; ref_out.backward(dout)      (float16 rounding points kept exactly)

Run from the repo root:  python oracle/make_golden.py
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def sdpa_case(seed, B, H, N, d, causal=False, scale=1.0):
    torch.manual_seed(seed)
    Q = torch.randn(B, H, N, d, dtype=torch.float32, requires_grad=True)
    K = torch.randn(B, H, N, d, dtype=torch.float32, requires_grad=True)
    V = torch.randn(B, H, N, d, dtype=torch.float32, requires_grad=True)
    O = torch.nn.functional.scaled_dot_product_attention(Q, K, V, scale=scale, is_causal=causal)
    dO = torch.randn(B, H, N, d, dtype=torch.float32)
    dQ, dK, dV = torch.autograd.grad(O, (Q, K, V), dO)
    lse = torch.logsumexp(scale * (Q.detach().double() @ K.detach().double().transpose(-1, -2))
                          .masked_fill(~torch.ones(N, N, dtype=torch.bool).tril() if causal
                                       else torch.zeros(N, N, dtype=torch.bool), float("-inf")), -1)
    return dict(Q=Q.detach().numpy(), K=K.detach().numpy(), V=V.detach().numpy(), dO=dO.numpy(),
                O=O.detach().numpy(), dQ=dQ.numpy(), dK=dK.numpy(), dV=dV.numpy(),
                lse=lse.float().numpy(), causal=np.array(causal), scale=np.array(scale, np.float32),
                seed=np.array(seed))


def tutorial_case(Z=1, H=2, N_CTX=128, HEAD_DIM=64, causal=True, dtype=torch.float16):
    torch.manual_seed(20)
    q = torch.empty((Z, H, N_CTX, HEAD_DIM), dtype=dtype).normal_(mean=0.0, std=0.5).requires_grad_()
    k = torch.empty((Z, H, N_CTX, HEAD_DIM), dtype=dtype).normal_(mean=0.0, std=0.5).requires_grad_()
    v = torch.empty((Z, H, N_CTX, HEAD_DIM), dtype=dtype).normal_(mean=0.0, std=0.5).requires_grad_()
    sm_scale = 0.5
    dout = torch.randn_like(q)
    M = torch.tril(torch.ones((N_CTX, N_CTX)))
    p = torch.matmul(q, k.transpose(2, 3)) * sm_scale
    if causal:
        p[:, :, M == 0] = float("-inf")
    p = torch.softmax(p.float(), dim=-1).to(dtype)
    ref_out = torch.matmul(p, v)
    ref_out.backward(dout)
    f = lambda t: t.detach().float().numpy()
    return dict(Q=f(q), K=f(k), V=f(v), dO=f(dout), O=f(ref_out), dQ=f(q.grad), dK=f(k.grad), dV=f(v.grad),
                causal=np.array(causal), scale=np.array(sm_scale, np.float32), seed=np.array(20))


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)  # keep the reduction order of the fixtures independent of the host
    cases = {
        # the reference's distribution and scale (test_correctness.py), seeds 0 and 1, shrunk shapes
        "sdpa_seed0_B1H2N64d16": sdpa_case(0, 1, 2, 64, 16),
        "sdpa_seed1_B2H2N96d32": sdpa_case(1, 2, 2, 96, 32),
        # BASELINE config 1 at reduced N (full C1 is run live by the tests): fp32 H=4 d=64 non-causal
        "sdpa_seed2_B1H4N128d64": sdpa_case(2, 1, 4, 128, 64),
        # causal + scale through the same SDPA call
        "sdpa_seed3_B1H2N128d64_causal": sdpa_case(3, 1, 2, 128, 64, causal=True, scale=0.125),
        # gradcheck shape of test_torch.py:4-13 (B=2,H=2,N=32,d=128, seed 5)
        "sdpa_seed5_B2H2N32d128": sdpa_case(5, 2, 2, 32, 128),
        # the vendored tutorial's causal fp16 check at reduced N_CTX
        "tutorial_seed20_Z1H2N128d64_causal": tutorial_case(),
    }
    for name, arrs in cases.items():
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrs)
        print(name, {k: v.shape for k, v in arrs.items() if hasattr(v, "shape") and v.ndim > 0})
    return 0


if __name__ == "__main__":
    sys.exit(main())
