"""GPU parity tests (B200): the CUDA path, called through the C ABI, against the CPU oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star; reference src/test_correctness.py:40,60-62):
  float32          O, L : max-abs <= 1e-4 ;  dQ 9e-4, dK 7e-4, dV 7e-5 (atol, rtol 1e-5) on the reference's own case
  float16/bfloat16 O, L : max-abs <= 2e-3 against the fp32/fp64 ground truth (inputs rounded to the kernel dtype first).
                   Two quantisation steps of the SPECIFIED arithmetic can exceed 2e-3 on their own and
                   are allowed on top, entry by entry, with their exact first-order bounds:
                     * the output dtype's rounding: half_ulp(|O_ref|)  (bf16: up to 2^-8 |O|, e.g. causal rows that see
                       one or two keys have |O| ~ |V|);
                     * P cast to the input dtype before P.V, as the reference does (flash_attention_kernels.py:98):
                       2^-(mant+2) * (P |V|)  (bf16: 2^-9 sum_j P_ij |V_jd|).
                   The plain 2e-3 is asserted wherever it is attainable: non-causal float16 at any scale; non-causal
                   bfloat16 at N >= 1024 with softmax_scale <= 1/sqrt(D), and on the tutorial distribution.  The measured
                   maxima per BASELINE config are in profiles/r02_parity_errors.json (tools/parity_errors.py).
  gradients        max|g - g_ref| / max|g_ref| <= 1e-2
  backward         bit-identical across repeated runs
"""
import glob
import math
import os

import numpy as np
import pytest
import torch

from flash_attention_dlrs_b200 import (FlashAttention, FlashAttentionDeterministic, _lib, _native,
                                       flash_attention_backward, flash_attention_forward)
from oracle import attention_oracle as orc

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0)
MANT_BITS = {torch.float16: 10, torch.bfloat16: 7}


def out_half_ulp(ref: torch.Tensor, dtype) -> torch.Tensor:
    """Exact bound of the output dtype's own rounding error at each reference value: 2^(floor(log2|x|) - mant - 1)."""
    if dtype == torch.float32:
        return torch.zeros_like(ref)
    e = torch.floor(torch.log2(ref.abs().clamp_min(2.0 ** -14)))
    return torch.exp2(e - MANT_BITS[dtype] - 1)


def make_inputs(seed, B, H, N, D, dtype, dist="randn"):
    """Seeded like the reference (torch.manual_seed(seed); randn fp32 then cast, test_correctness.py:29-32,46)."""
    g = torch.Generator().manual_seed(seed)
    std = 0.5 if dist == "tutorial" else 1.0
    t = [(torch.randn(B, H, N, D, generator=g) * std).to(dtype) for _ in range(4)]
    return t  # Q, K, V, dO on CPU in the kernel dtype


def run_gpu(Q, K, V, dO, causal, scale):
    q, k, v, do = (t.to(DEV) for t in (Q, K, V, dO))
    O, L = flash_attention_forward(q, k, v, DEV, causal, scale)
    dQ, dK, dV = flash_attention_backward(q, k, v, O, do, L, DEV, True, causal, scale)
    torch.cuda.synchronize()
    return O.cpu(), L.cpu(), dQ.cpu(), dK.cpu(), dV.cpu()


def rel_err(a, ref):
    """max|g - g_ref| / max|g_ref| ; a reference gradient that is identically ~0 (N = 1: dS = 0) is compared absolutely."""
    return ((a.double() - ref.double()).abs().max() / ref.double().abs().max().clamp_min(1e-3)).item()


def o_bound(Q, K, V, O_ref, scale, causal, dtype):
    """2e-3 + output half-ulp + first-order bound of rounding P to the input dtype (module docstring)."""
    if dtype == torch.float32:
        return torch.full_like(O_ref, 1e-4)
    p_absv = orc.reference_sdpa(Q.float(), K.float(), V.float().abs(), scale, causal).to(O_ref.dtype)
    return 2e-3 + out_half_ulp(O_ref, dtype) + 2.0 ** -(MANT_BITS[dtype] + 2) * p_absv


def check_case(seed, B, H, N, D, dtype, causal, scale, dist="randn", strict_o=None):
    Q, K, V, dO = make_inputs(seed, B, H, N, D, dtype, dist)
    O, L, dQ, dK, dV = run_gpu(Q, K, V, dO, causal, scale)
    assert O.dtype == dtype and L.dtype == torch.float32 and L.shape == (B, H, N, 1)
    ref = orc.attention_grads_fp64(Q.float(), K.float(), V.float(), dO.float(), scale, causal)
    o_err = (O.double() - ref["O"]).abs()
    l_err = (L.double() - ref["L"]).abs().max().item()
    if dtype == torch.float32:
        assert o_err.max().item() <= 1e-4, f"O err {o_err.max().item():.3e}"
        assert l_err <= 1e-4, f"L err {l_err:.3e}"
    else:
        bound = o_bound(Q, K, V, ref["O"], scale, causal, dtype)
        assert (o_err <= bound).all(), f"O err {o_err.max().item():.3e}"
        # the north-star's plain 2e-3 wherever it is attainable: non-causal float16 everywhere; non-causal bfloat16 where
        # rows average over enough keys that |O| < 0.5, i.e. the output's own half-ulp is < 1e-3 (N >= 1024 at a softmax
        # scale <= 1/sqrt(D), SURVEY.md section 0-10: floor 1.0e-3; measured 2.3e-3 at N = 128, where |O| reaches 0.5-1
        # and the bf16 half-ulp alone is 1.95e-3) and on the tutorial distribution (|V| ~ 0.5).  Causal rows that see one
        # or two keys have |O| ~ |V| and keep the entry-wise `bound`.
        attainable = not causal and N >= 64 and (
            dtype == torch.float16 or dist == "tutorial" or (N >= 1024 and scale * math.sqrt(D) <= 1.05))
        if strict_o if strict_o is not None else attainable:
            assert o_err.max().item() <= 2e-3, f"O err {o_err.max().item():.3e} (strict)"
        assert l_err <= 2e-3, f"L err {l_err:.3e}"
    for name, got in (("dQ", dQ), ("dK", dK), ("dV", dV)):
        e = rel_err(got, ref[name])
        assert e <= 1e-2, f"{name} rel err {e:.3e}"
    return O, L, dQ, dK, dV


# ------------------------------------------------------------------------------------------------ float32
def _golden(golden_dir, prefix):
    files = sorted(glob.glob(os.path.join(golden_dir, prefix + "*.npz")))
    assert files
    return files


@pytest.mark.parametrize("idx", range(5))
def test_fp32_golden_vectors_reference_tolerances(golden_dir, idx):
    """Committed ground truth of the reference's own check, at its own tolerances (test_correctness.py:40,60-62)."""
    z = np.load(_golden(golden_dir, "sdpa_")[idx])
    t = {k: torch.from_numpy(z[k]) for k in z.files if z[k].ndim > 0}
    causal, scale = bool(z["causal"]), float(z["scale"])
    O, L, dQ, dK, dV = run_gpu(t["Q"], t["K"], t["V"], t["dO"], causal, scale)
    assert torch.allclose(O, t["O"], atol=1e-4, rtol=1e-5)
    assert torch.allclose(dQ, t["dQ"], atol=9e-4, rtol=1e-5)
    assert torch.allclose(dK, t["dK"], atol=7e-4, rtol=1e-5)
    assert torch.allclose(dV, t["dV"], atol=7e-5, rtol=1e-5)
    assert torch.allclose(L.squeeze(-1) * math.log(2.0), t["lse"], atol=1e-4, rtol=1e-5)


def test_fp32_baseline_config1():
    """BASELINE.json configs[0]: fp32 B=1 H=4 N=512 D=64 non-causal, scale 1 — vs the reference's CPU ground truth."""
    Q, K, V, dO = make_inputs(0, 1, 4, 512, 64, torch.float32)
    O, L, dQ, dK, dV = run_gpu(Q, K, V, dO, False, 1.0)
    O_ref, dQ_ref, dK_ref, dV_ref = orc.reference_sdpa_grads(Q, K, V, dO, 1.0, False)
    assert torch.allclose(O, O_ref, atol=1e-4, rtol=1e-5)
    assert torch.allclose(dQ, dQ_ref, atol=9e-4, rtol=1e-5)
    assert torch.allclose(dK, dK_ref, atol=7e-4, rtol=1e-5)
    assert torch.allclose(dV, dV_ref, atol=7e-5, rtol=1e-5)


@pytest.mark.parametrize("seed", [0, 1])
def test_fp32_reference_correctness_shape(seed):
    """The reference's test shape (test_correctness.py:9-14: B=32,H=32,N=256,d=128, randn, scale 1) on a batch slice."""
    Q, K, V, dO = make_inputs(seed, 4, 32, 256, 128, torch.float32)
    O, L, dQ, dK, dV = run_gpu(Q, K, V, dO, False, 1.0)
    O_ref, dQ_ref, dK_ref, dV_ref = orc.reference_sdpa_grads(Q, K, V, dO, 1.0, False)
    assert torch.allclose(O, O_ref, atol=1e-4, rtol=1e-5)
    assert torch.allclose(dQ, dQ_ref, atol=9e-4, rtol=1e-5)
    assert torch.allclose(dK, dK_ref, atol=7e-4, rtol=1e-5)
    assert torch.allclose(dV, dV_ref, atol=7e-5, rtol=1e-5)


@pytest.mark.parametrize("N,D,causal,scale", [(96, 16, True, 0.25), (200, 32, False, 1.0), (333, 64, True, 0.125),
                                              (64, 128, True, 1.0), (1, 16, False, 1.0), (130, 40, True, 0.2)])
def test_fp32_ragged_and_padded(N, D, causal, scale):
    check_case(3, 2, 3, N, D, torch.float32, causal, scale)


def test_fp32_gradcheck_like_reference():
    """src/test_torch.py:4-13,30: gradcheck, fp32, B=2,H=2,N=32,d=128, seed 5, eps 2e-2, atol=rtol=1e-2.
    The reference allows nondet_tol=1e-4; the deterministic backward passes with 0."""
    torch.manual_seed(5)
    Q = torch.randn(2, 2, 32, 128, dtype=torch.float32, device=DEV, requires_grad=True)
    K = torch.randn(2, 2, 32, 128, dtype=torch.float32, device=DEV, requires_grad=True)
    V = torch.randn(2, 2, 32, 128, dtype=torch.float32, device=DEV, requires_grad=True)
    for fn in (FlashAttention.apply, FlashAttentionDeterministic.apply):
        assert torch.autograd.gradcheck(fn, (Q, K, V), eps=2e-2, atol=1e-2, rtol=1e-2, nondet_tol=0.0)


# ------------------------------------------------------------------------------------------------ 16-bit tcgen05 path
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("D", [64, 128])
@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("N", [128, 256, 384, 1024])
def test_16bit_parity(dtype, D, causal, N):
    check_case(N + D, 2, 3, N, D, dtype, causal, 1.0 / math.sqrt(D))


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("causal", [False, True])
def test_16bit_tutorial_distribution(dtype, causal):
    """normal(0, 0.5), sm_scale 0.5 (flash_attention_openai_tutorial.py:527-530)."""
    check_case(20, 1, 2, 1024, 64, dtype, causal, 0.5, dist="tutorial")


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("N", [1, 100, 129, 200, 700])
@pytest.mark.parametrize("causal", [False, True])
def test_16bit_ragged_sequence(dtype, N, causal):
    check_case(N, 1, 2, N, 64, dtype, causal, 0.125)
    check_case(N + 1, 1, 2, N, 128, dtype, causal, 0.09)


@pytest.mark.parametrize("dtype,d", [(torch.bfloat16, 80), (torch.float16, 40), (torch.bfloat16, 16)])
def test_16bit_padded_head_size(dtype, d):
    """Head sizes that need padding — forward AND backward (the reference's padded backward is broken)."""
    check_case(9, 1, 2, 256, d, dtype, True, 1.0 / math.sqrt(d))


def test_16bit_reference_default_scale_one():
    """scale = 1, randn (the reference's default): softmax is nearly one-hot and |O| ~ |V|, so O is judged against
    the output dtype's rounding floor; L must still be within 2e-3... of a logit of size ~50,
    so L is compared relatively."""
    Q, K, V, dO = make_inputs(1, 1, 2, 512, 128, torch.float16)
    O, L, dQ, dK, dV = run_gpu(Q, K, V, dO, False, 1.0)
    ref = orc.attention_grads_fp64(Q.float(), K.float(), V.float(), dO.float(), 1.0, False)
    assert ((O.double() - ref["O"]).abs() <= o_bound(Q, K, V, ref["O"], 1.0, False, torch.float16)).all()
    assert ((L.double() - ref["L"]).abs() <= 2e-3).all()
    for name, got in (("dQ", dQ), ("dK", dK), ("dV", dV)):
        assert rel_err(got, ref[name]) <= 1e-2, name


def test_fp16_tutorial_golden(golden_dir):
    """The vendored tutorial's own check (flash_attention_openai_tutorial.py:551-559): atol 1e-2, rtol 0."""
    z = np.load(_golden(golden_dir, "tutorial_")[0])
    t = {k: torch.from_numpy(z[k]).to(torch.float16) for k in ("Q", "K", "V", "dO")}
    O, L, dQ, dK, dV = run_gpu(t["Q"], t["K"], t["V"], t["dO"], True, 0.5)
    for got, key in ((O, "O"), (dQ, "dQ"), (dK, "dK"), (dV, "dV")):
        assert torch.allclose(got.float(), torch.from_numpy(z[key]), atol=1e-2, rtol=0), key


# ------------------------------------------------------------------------------------------------ determinism / views
@pytest.mark.parametrize("dtype,N,D", [(torch.bfloat16, 1024, 128), (torch.float16, 640, 64), (torch.float32, 200, 64)])
def test_backward_bit_identical_across_runs(dtype, N, D):
    Q, K, V, dO = (t.to(DEV) for t in make_inputs(4, 2, 4, N, D, dtype))
    O, L = flash_attention_forward(Q, K, V, DEV, True, 0.1)
    first = flash_attention_backward(Q, K, V, O, dO, L, DEV, False, True, 0.1)
    for _ in range(10):
        again = flash_attention_backward(Q, K, V, O, dO, L, DEV, False, True, 0.1)
        for a, b in zip(first, again):
            assert torch.equal(a, b)
    O2, L2 = flash_attention_forward(Q, K, V, DEV, True, 0.1)
    assert torch.equal(O, O2) and torch.equal(L, L2)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_head_slice_views_match_full_run_bitwise(dtype):
    """Head-sharded use: a strided view X[:, h0:h1] gives exactly the bits of the same heads in the full run."""
    Q, K, V, dO = (t.to(DEV) for t in make_inputs(6, 2, 6, 256, 64, dtype))
    O, L = flash_attention_forward(Q, K, V, DEV, True, 0.125)
    g = flash_attention_backward(Q, K, V, O, dO, L, DEV, False, True, 0.125)
    sl = slice(2, 5)
    Os, Ls = flash_attention_forward(Q[:, sl], K[:, sl], V[:, sl], DEV, True, 0.125)
    gs = flash_attention_backward(Q[:, sl], K[:, sl], V[:, sl], Os, dO[:, sl], Ls, DEV, False, True, 0.125)
    assert torch.equal(Os, O[:, sl]) and torch.equal(Ls, L[:, sl])
    for a, b in zip(gs, g):
        assert torch.equal(a, b[:, sl])


def test_autograd_function_matches_functional_pair():
    Q, K, V, dO = (t.to(DEV) for t in make_inputs(8, 1, 2, 512, 128, torch.bfloat16))
    q, k, v = (t.clone().requires_grad_(True) for t in (Q, K, V))
    O = FlashAttention.apply(q, k, v, True, 0.09)
    O.backward(dO)
    O2, L2 = flash_attention_forward(Q, K, V, DEV, True, 0.09)
    dQ, dK, dV = flash_attention_backward(Q, K, V, O2, dO, L2, DEV, False, True, 0.09)
    assert torch.equal(O.detach(), O2)
    assert torch.equal(q.grad, dQ) and torch.equal(k.grad, dK) and torch.equal(v.grad, dV)
    # reference call form: no extra arguments -> scale 1.0, non-causal
    O3 = FlashAttention.apply(Q, K, V)
    O4, _ = flash_attention_forward(Q, K, V, DEV)
    assert torch.equal(O3, O4)


@pytest.mark.parametrize("dtype,D,N", [(torch.bfloat16, 128, 512), (torch.float16, 64, 333), (torch.float32, 32, 200),
                                       (torch.bfloat16, 64, 1)])
@pytest.mark.parametrize("causal", [False, True])
def test_cpp_autograd_node_is_bitwise_the_python_function(dtype, D, N, causal, monkeypatch):
    """FlashAttention.apply's plain form runs the C++ autograd node (csrc/torch_binding.cpp); forcing the Python Function
    must give the same bits, for contiguous inputs and for strided views (a head slice, a transposed layout)."""
    from flash_attention_dlrs_b200 import flash_attention_torch as fat
    node = fat._cpp_node()
    assert node is not None, "the torch binding is not built"
    base = [t.to(DEV) for t in make_inputs(21, 2, 5, N, D, dtype)]
    views = [[t[:, 1:4] for t in base],                                            # head slice: strided, kernel-legal
             [t.transpose(1, 2).contiguous().transpose(1, 2) for t in base]]       # (B, N, H, D) memory layout
    for Q, K, V, dO in views:
        assert node.supported(Q, K, V)
        outs = []
        for use_cpp in (True, False):
            monkeypatch.setattr(fat, "USE_CPP_NODE", use_cpp)
            q, k, v = (t.detach().clone().requires_grad_(True) if False else t.detach().requires_grad_(True) for t in (Q, K, V))
            O = FlashAttention.apply(q, k, v, causal, 0.11)
            assert (type(O.grad_fn).__name__ == "FlashAttentionBackward") != use_cpp   # which node recorded the graph
            O.backward(dO)
            outs.append((O.detach(), q.grad, k.grad, v.grad))
        for a, b in zip(*outs):
            assert torch.equal(a, b)
    # no graph under no_grad, and unsupported forms fall through to the Python Function (padded head size)
    with torch.no_grad():
        assert FlashAttention.apply(*views[0][:3], causal, 0.11).grad_fn is None
    Qp = torch.randn(1, 2, 64, 40, device=DEV, dtype=torch.bfloat16)
    assert not node.supported(Qp, Qp, Qp) and FlashAttention.apply(Qp, Qp, Qp).shape == Qp.shape


@pytest.mark.parametrize("dtype,D", [(torch.bfloat16, 128), (torch.float16, 64)])
@pytest.mark.parametrize("Nq,Nk", [(256, 512), (512, 256), (200, 333), (1, 130), (384, 128)])
def test_rectangular_attention_matches_the_oracle(dtype, D, Nq, Nk):
    """fa_fwd_rect / fa_bwd_rect: Nq query rows against Nk key rows (the ring path's one-chunk-against-a-shard step, and
    cross-attention) against the fp64 closed form; Nk == Nq must reproduce the square entry points bit for bit."""
    B, H, scale = 2, 3, 1.0 / math.sqrt(D)
    g = torch.Generator().manual_seed(5)
    Q, dO = (torch.randn(B, H, Nq, D, generator=g).to(dtype).to(DEV) for _ in range(2))
    K, V = (torch.randn(B, H, Nk, D, generator=g).to(dtype).to(DEV) for _ in range(2))
    O, L = _native.forward_rect(Q, K, V, scale)
    dQ, dK, dV = _native.backward_rect(Q, K, V, O, dO, L, scale)
    q, k, v, do = (t.cpu().double() for t in (Q, K, V, dO))
    q.requires_grad_(True), k.requires_grad_(True), v.requires_grad_(True)
    S = (q @ k.transpose(-1, -2)) * scale
    ref_O = torch.softmax(S, -1) @ v
    ref_L = torch.logsumexp(S, -1) * math.log2(math.e)
    gq, gk, gv = torch.autograd.grad(ref_O, (q, k, v), do)
    assert (L.cpu().double() - ref_L).abs().max() <= 2e-3
    assert (O.cpu().double() - ref_O).abs().max() <= 2e-3 + 2.0 ** -8 * ref_O.abs().max()
    for name, got, ref in (("dQ", dQ, gq), ("dK", dK, gk), ("dV", dV, gv)):
        assert rel_err(got.cpu(), ref) <= 1e-2, name
    # square problems through the rectangular entry points are the square kernels
    K2, V2 = (torch.randn(B, H, Nq, D, generator=g).to(dtype).to(DEV) for _ in range(2))
    O2, L2 = _native.forward_rect(Q, K2, V2, scale)
    O3, L3 = _native.forward(Q, K2, V2, False, scale)
    assert torch.equal(O2, O3) and torch.equal(L2, L3)
    for a, b in zip(_native.backward_rect(Q, K2, V2, O2, dO, L2, scale), _native.backward(Q, K2, V2, O3, dO, L3, False, scale)):
        assert torch.equal(a, b)
    # what the rectangular path does not take
    lib = _lib.load()
    s4 = _lib.strides4(Q)
    assert lib.fa_fwd_rect(_native._ptr(Q), _native._ptr(K), _native._ptr(V), _native._ptr(O), _native._ptr(L), B, H, Nq, 0, D,
                           s4, _lib.strides4(K), _lib.strides4(V), s4, _native.dtype_code(dtype), scale, None) < 0


def test_preprocess_parity():
    for dtype, D in ((torch.bfloat16, 128), (torch.float16, 64), (torch.float32, 32)):
        O, dO = (torch.randn(2, 5, 300, D, device=DEV).to(dtype) for _ in range(2))
        delta = _native.backward_preprocess(O[:, 1:4], dO[:, 1:4])
        ref = (O[:, 1:4].double() * dO[:, 1:4].double()).sum(-1)
        assert (delta.double() - ref).abs().max().item() <= 1e-4 * max(1.0, ref.abs().max().item())


def test_errors_on_gpu_tensors():
    Q = torch.randn(1, 1, 16, 16, device=DEV, dtype=torch.float64)
    with pytest.raises(TypeError, match="not supported"):
        FlashAttention.apply(Q, Q, Q)
    Qf = torch.randn(1, 1, 16, 16, device=DEV)
    with pytest.raises(ValueError):
        FlashAttention.apply(Qf, Qf[:, :, :8], Qf)
    with pytest.raises(ValueError):
        FlashAttention.apply(Qf, Qf.half(), Qf)
    with pytest.raises(ValueError):
        FlashAttention.apply(Qf[0], Qf[0], Qf[0])
    with pytest.raises(_lib.FlashAttentionLibraryError):
        flash_attention_forward(Qf, Qf, Qf, DEV, False, -1.0)


# ------------------------------------------------------------------------------------------------ BASELINE full sizes
def _subset_oracle_check(Q, K, V, dO, O, L, grads, heads, causal, scale, dtype):
    for (b, h) in heads:
        sl = (slice(b, b + 1), slice(h, h + 1))
        q, k, v, do = (t[sl].float().cpu() for t in (Q, K, V, dO))
        O_ref, dQ_ref, dK_ref, dV_ref = orc.reference_sdpa_grads(q, k, v, do, scale, causal)
        o_err = (O[sl].float().cpu() - O_ref).abs()
        assert (o_err <= o_bound(q, k, v, O_ref, scale, causal, dtype)).all(), f"O err {o_err.max().item():.3e} head {(b, h)}"
        for name, got, ref in (("dQ", grads[0], dQ_ref), ("dK", grads[1], dK_ref), ("dV", grads[2], dV_ref)):
            assert rel_err(got[sl].float().cpu(), ref) <= 1e-2, f"{name} head {(b, h)}"


def test_config2_full_size_fp16():
    """BASELINE configs[1]: fwd fp16 B=4 H=16 N=4096 D=64 non-causal.  Full-size run; two heads against the CPU
    ground truth, plus size-independent properties: softmax rows sum to one (V = 1 -> O = 1) and linearity in V."""
    B, H, N, D, scale = 4, 16, 4096, 64, 0.125
    Q, K, V, dO = (t.to(DEV) for t in make_inputs(42, B, H, N, D, torch.float16))
    O, L = flash_attention_forward(Q, K, V, DEV, False, scale)
    g = flash_attention_backward(Q, K, V, O, dO, L, DEV, False, False, scale)
    _subset_oracle_check(Q, K, V, dO, O, L, g, [(0, 0), (3, 15)], False, scale, torch.float16)
    ones = torch.ones_like(V)
    O1, L1 = flash_attention_forward(Q, K, ones, DEV, False, scale)
    assert (O1.float() - 1.0).abs().max().item() <= 2e-3
    assert torch.equal(L1, L)
    O2, _ = flash_attention_forward(Q, K, 2 * V, DEV, False, scale)
    assert (O2.float() - 2 * O.float()).abs().max().item() <= 4e-3


def test_config3_full_size_bf16_causal():
    """BASELINE configs[2]: fwd+bwd bf16 B=2 H=32 N=8192 D=128 causal with deterministic backward."""
    B, H, N, D = 2, 32, 8192, 128
    scale = 1.0 / math.sqrt(D)
    Q, K, V, dO = (t.to(DEV) for t in make_inputs(42, B, H, N, D, torch.bfloat16))
    O, L = flash_attention_forward(Q, K, V, DEV, True, scale)
    g = flash_attention_backward(Q, K, V, O, dO, L, DEV, False, True, scale)
    g2 = flash_attention_backward(Q, K, V, O, dO, L, DEV, False, True, scale)
    for a, b in zip(g, g2):
        assert torch.equal(a, b)                       # bit-identical backward
    assert all(torch.isfinite(t.float()).all() for t in (O, L, *g))
    _subset_oracle_check(Q, K, V, dO, O, L, g, [(0, 0), (1, 31)], True, scale, torch.bfloat16)
    # causal: row 0 attends only to key 0 -> O[0] = V[0] exactly, L[0] = log2e * scale * q0.k0
    assert torch.equal(O[:, :, 0], V[:, :, 0])
    l0 = (Q[:, :, 0].float() * K[:, :, 0].float()).sum(-1) * scale * orc.LOG2_E
    assert (L[:, :, 0, 0] - l0).abs().max().item() <= 1e-3
    # head-sharded slice == full run, bit for bit
    Os, Ls = flash_attention_forward(Q[:, 8:12], K[:, 8:12], V[:, 8:12], DEV, True, scale)
    assert torch.equal(Os, O[:, 8:12]) and torch.equal(Ls, L[:, 8:12])


def test_config4_long_context_full_size():
    """BASELINE configs[3]: long-context fwd+bwd bf16 B=1 H=64 N=32768 D=128 causal (the head-sharded config, here all 64
    heads on one GPU).  The N x N problem is too big for a CPU oracle per head, so: exact closed-form oracle on a subset
    of query rows (O, L, dQ rows only need one row of P), checksum properties for dK / dV, bitwise determinism, and
    bit-equality of a head-sharded slice with the full run."""
    B, H, N, D = 1, 64, 32768, 128
    scale = 1.0 / math.sqrt(D)
    g = torch.Generator(device=DEV).manual_seed(1234)
    Q, K, V, dO = (torch.randn(B, H, N, D, device=DEV, generator=g).to(torch.bfloat16) for _ in range(4))
    O, L = flash_attention_forward(Q, K, V, DEV, True, scale)
    dQ, dK, dV = flash_attention_backward(Q, K, V, O, dO, L, DEV, False, True, scale)
    dQ2, dK2, dV2 = flash_attention_backward(Q, K, V, O, dO, L, DEV, False, True, scale)
    assert torch.equal(dQ, dQ2) and torch.equal(dK, dK2) and torch.equal(dV, dV2)
    assert all(torch.isfinite(t.float()).all() for t in (O, L, dQ, dK, dV))

    # (1) row-subset oracle in float64 on the CPU: rows of P need only q_i and all keys <= i
    rows = [0, 1, 127, 128, 4097, 16383, 20000, 32767]
    dq_head_max = dQ[0].float().abs().amax(dim=(1, 2)).double().cpu()
    for h in (0, 37, 63):
        k64, v64 = K[0, h].double().cpu(), V[0, h].double().cpu()
        for i in rows:
            q = Q[0, h, i].double().cpu()
            s = scale * (k64[: i + 1] @ q)
            lse = torch.logsumexp(s, 0)
            p = torch.exp(s - lse)
            o_ref = p @ v64[: i + 1]
            do = dO[0, h, i].double().cpu()
            o_err = (O[0, h, i].double().cpu() - o_ref).abs()
            bound = 2e-3 + out_half_ulp(o_ref, torch.bfloat16) + 2.0 ** -9 * (p @ v64[: i + 1].abs())
            assert (o_err <= bound).all(), f"O row {i} head {h}: {o_err.max().item():.3e}"
            assert abs(L[0, h, i, 0].item() - lse.item() * orc.LOG2_E) <= 2e-3
            dp = v64[: i + 1] @ do
            ds = p * (dp - (o_ref * do).sum())
            dq_ref = scale * (ds @ k64[: i + 1])
            # same normalisation as everywhere else: the tensor-level max |dQ| (of this head), not the row's own
            e = (dQ[0, h, i].double().cpu() - dq_ref).abs().max() / dq_head_max[h]
            assert e.item() <= 1e-2, f"dQ row {i} head {h}: {e.item():.3e}"

    # (2) checksums: rows of P sum to one  =>  sum_j dV_j = sum_i dO_i ;  rows of dS sum to zero  =>  sum_j dK_j = 0
    sum_dv = dV.float().sum(2)
    sum_do = dO.float().sum(2)
    assert ((sum_dv - sum_do).abs().max() / sum_do.abs().max()).item() <= 1e-2
    dk_scale = dK.float().abs().sum(2).max().item()
    assert (dK.float().sum(2).abs().max().item() / dk_scale) <= 1e-2

    # (3) head-sharded slice (rank 3 of 8 owns heads 24..31) is bit-identical to the full run
    sl = slice(24, 32)
    Os, Ls = flash_attention_forward(Q[:, sl], K[:, sl], V[:, sl], DEV, True, scale)
    gs = flash_attention_backward(Q[:, sl], K[:, sl], V[:, sl], Os, dO[:, sl], Ls, DEV, False, True, scale)
    assert torch.equal(Os, O[:, sl]) and torch.equal(Ls, L[:, sl])
    for a, b in zip(gs, (dQ, dK, dV)):
        assert torch.equal(a, b[:, sl])


def test_host_pipeline_matches_resident_path_bitwise():
    """attention_from_host (pinned host tensors, chunked H2D / compute / D2H pipeline) == functional pair, bit for bit."""
    from flash_attention_dlrs_b200 import attention_from_host

    host = [t.pin_memory() for t in make_inputs(12, 2, 6, 384, 128, torch.bfloat16)]
    O, dQ, dK, dV = attention_from_host(*host, causal=True, softmax_scale=0.09, device=DEV, chunks=4)
    Q, K, V, dO = (t.to(DEV) for t in host)
    O2, L2 = flash_attention_forward(Q, K, V, DEV, True, 0.09)
    g = flash_attention_backward(Q, K, V, O2, dO, L2, DEV, False, True, 0.09)
    assert torch.equal(O, O2.cpu())
    for a, b in zip((dQ, dK, dV), g):
        assert torch.equal(a, b.cpu())
    O3 = attention_from_host(*host[:3], causal=False, softmax_scale=0.09, device=DEV, chunks=5)
    assert torch.equal(O3, flash_attention_forward(Q, K, V, DEV, False, 0.09)[0].cpu())
    with pytest.raises(ValueError):
        attention_from_host(Q, K, V)


# ------------------------------------------------------------------------------------------------ backward variants
@pytest.mark.parametrize("dtype,D", [(torch.bfloat16, 128), (torch.float16, 64)])
@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("N", [128, 200, 640, 1536])
def test_two_kernel_backward_matches_oracle_and_fused(dtype, D, causal, N):
    """The two backward implementations — fa_bwd's two-kernel path and the single-pass kernel with the ordered dQ
    reduction (FA_BWD_FUSED) — against the oracle.  Under the causal mask both visit the query blocks of a key block in
    the same order, so dK and dV agree bit for bit."""
    B, H, scale = 2, 3, 1.0 / math.sqrt(D)
    Q, K, V, dO = (t.to(DEV) for t in make_inputs(11, B, H, N, D, dtype))
    O, L = flash_attention_forward(Q, K, V, DEV, causal, scale)
    two = _native.backward(Q, K, V, O, dO, L, causal, scale, which=_native.BWD_DKDV | _native.BWD_DQ)
    one = _native.backward(Q, K, V, O, dO, L, causal, scale, which=_native.BWD_FUSED)
    torch.cuda.synchronize()
    ref = orc.attention_grads_fp64(Q.cpu().float(), K.cpu().float(), V.cpu().float(), dO.cpu().float(), scale, causal)
    for name, a, b in zip(("dQ", "dK", "dV"), two, one):
        assert rel_err(a.cpu(), ref[name]) <= 1e-2, name
        assert rel_err(b.cpu(), ref[name]) <= 1e-2, name
    if causal:
        assert torch.equal(two[1], one[1]) and torch.equal(two[2], one[2])


@pytest.mark.parametrize("causal", [False, True])
def test_backward_variants_bit_identical_many_heads(causal):
    """Ordered dQ reduction: more CTAs than SMs, so turn-taking between waves is exercised; 5 runs must agree bitwise."""
    B, H, N, D = 2, 24, 1024, 128
    Q, K, V, dO = (t.to(DEV) for t in make_inputs(3, B, H, N, D, torch.bfloat16))
    O, L = flash_attention_forward(Q, K, V, DEV, causal, 0.09)
    for which in (_native.BWD_FUSED, _native.BWD_DKDV | _native.BWD_DQ):
        first = _native.backward(Q, K, V, O, dO, L, causal, 0.09, which=which)
        for _ in range(4):
            again = _native.backward(Q, K, V, O, dO, L, causal, 0.09, which=which)
            for a, b in zip(first, again):
                assert torch.equal(a, b)


# ------------------------------------------------------------------------------------------------ streams and graphs
def test_cuda_graph_capture_and_replay_bitwise():
    """The C-ABI entry points never synchronise or allocate: forward + backward can be captured in a CUDA graph on a
    side stream and replayed; replays reproduce the eager result bit for bit (also for the single-pass backward, whose
    workspace counters are reset by a captured memset)."""
    B, H, N, D = 1, 4, 512, 128
    Q, K, V, dO = (t.to(DEV) for t in make_inputs(7, B, H, N, D, torch.bfloat16))
    eager_O, eager_L = _native.forward(Q, K, V, True, 0.09)
    eager_g = _native.backward(Q, K, V, eager_O, dO, eager_L, True, 0.09)
    eager_f = _native.backward(Q, K, V, eager_O, dO, eager_L, True, 0.09, which=_native.BWD_FUSED)
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):   # warm-up on the capture stream (function attributes, allocator pools)
            o, l = _native.forward(Q, K, V, True, 0.09)
            _native.backward(Q, K, V, o, dO, l, True, 0.09)
            _native.backward(Q, K, V, o, dO, l, True, 0.09, which=_native.BWD_FUSED)
    torch.cuda.current_stream().wait_stream(s)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        g_O, g_L = _native.forward(Q, K, V, True, 0.09)
        g_g = _native.backward(Q, K, V, g_O, dO, g_L, True, 0.09)
        g_f = _native.backward(Q, K, V, g_O, dO, g_L, True, 0.09, which=_native.BWD_FUSED)
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(g_O, eager_O) and torch.equal(g_L, eager_L)
    for a, b in zip(g_g + g_f, eager_g + eager_f):
        assert torch.equal(a, b)


def test_concurrent_streams_and_threads():
    """Re-entrant and thread-safe: two host threads launch on their own streams at the same time."""
    import threading
    B, H, N, D = 1, 4, 1024, 64
    Q, K, V, dO = (t.to(DEV) for t in make_inputs(9, B, H, N, D, torch.float16))
    want_O, want_L = _native.forward(Q, K, V, False, 0.125)
    want_g = _native.backward(Q, K, V, want_O, dO, want_L, False, 0.125)
    torch.cuda.synchronize()
    results, errors = {}, []

    def worker(i):
        try:
            st = torch.cuda.Stream()
            with torch.cuda.stream(st):
                for _ in range(5):
                    o, l = _native.forward(Q, K, V, False, 0.125)
                    g = _native.backward(Q, K, V, o, dO, l, False, 0.125)
            st.synchronize()
            results[i] = (o, l, g)
        except Exception as e:   # noqa: BLE001
            errors.append(e)

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for o, l, g in results.values():
        assert torch.equal(o, want_O) and torch.equal(l, want_L)
        for a, b in zip(g, want_g):
            assert torch.equal(a, b)


# ------------------------------------------------------------------------------------------------ FP8 forward
FP8_FORMATS = {torch.float8_e4m3fn: (3, -6), torch.float8_e5m2: (2, -14)}   # mantissa bits, smallest normal exponent


def fp8_o_bound(Qf, Kf, Vf, scale, causal, dtype):
    """Entry-wise bound of what the SPECIFIED FP8 arithmetic may differ from the fp64 ground truth by: 2e-3, plus the
    output's own rounding (half an ulp of the FP8 output, flash_attention_torch.py:50 allocates O in the input dtype),
    plus P rounded to FP8 before P.V (flash_attention_openai_tutorial.py:66-67: relative half-ulp 2^-(m+1), first order
    2^-(m+1) sum_j P_ij |V_jd|), plus the P entries below the smallest FP8 subnormal, which flush to zero."""
    m, emin = FP8_FORMATS[dtype]
    S = orc._scores_fp64(Qf, Kf, scale, causal)
    E = torch.exp(S - S.amax(-1, keepdim=True))
    P = E / E.sum(-1, keepdim=True)
    absV = Vf.double().abs()
    O = P @ Vf.double()
    half_ulp = torch.exp2(torch.floor(torch.log2(O.abs().clamp_min(2.0 ** emin))) - m - 1)
    p_round = 2.0 ** -(m + 1) * (P @ absV)
    flush = (P * (E < 2.0 ** (emin - m + 1))) @ absV
    return O, 2e-3 + half_ulp + p_round + flush


@pytest.mark.parametrize("dtype", list(FP8_FORMATS))
@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("N,d", [(128, 128), (384, 128), (1000, 128), (512, 64), (300, 80)])
def test_fp8_forward_parity(dtype, causal, N, d):
    B, H = 2, 3
    scale = 1.0 / math.sqrt(d)
    g = torch.Generator().manual_seed(21)
    Q, K, V = (torch.randn(B, H, N, d, generator=g).to(dtype) for _ in range(3))
    O, L = flash_attention_forward(Q.to(DEV), K.to(DEV), V.to(DEV), DEV, causal, scale)
    torch.cuda.synchronize()
    assert O.dtype == dtype and O.shape == (B, H, N, d) and L.dtype == torch.float32
    ref_O, bound = fp8_o_bound(Q.float(), K.float(), V.float(), scale, causal, dtype)
    err = (O.cpu().double() - ref_O).abs()
    assert (err <= bound).all(), f"O err {err.max().item():.3e} (worst excess {(err - bound).max().item():.3e})"
    ref = orc.attention_fp64(Q.float(), K.float(), V.float(), scale, causal)
    assert (L.cpu().double() - ref[1]).abs().max().item() <= 2e-3


def test_fp8_forward_with_key_padding_mask():
    """FP8 forward + per-batch valid lengths: each batch element equals the forward of its own first seqlens[b] tokens
    bit for bit (same tiles, same arithmetic); the padded rows are zero bytes."""
    B, H, N, D = 3, 2, 384, 128
    lens = [384, 130, 77]
    g = torch.Generator().manual_seed(8)
    Q, K, V = (torch.randn(B, H, N, D, generator=g).to(torch.float8_e4m3fn).to(DEV) for _ in range(3))
    O, L = flash_attention_forward(Q, K, V, DEV, True, 0.09, torch.tensor(lens))
    for b, n in enumerate(lens):
        Ob, Lb = flash_attention_forward(Q[b:b + 1, :, :n].contiguous(), K[b:b + 1, :, :n].contiguous(),
                                         V[b:b + 1, :, :n].contiguous(), DEV, True, 0.09)
        assert torch.equal(O[b:b + 1, :, :n].view(torch.uint8), Ob.view(torch.uint8)) and torch.equal(L[b:b + 1, :, :n], Lb)
        assert not O[b, :, n:].view(torch.uint8).any() and not L[b, :, n:].any()


def test_fp8_is_forward_only_and_deterministic():
    Q, K, V = (torch.randn(1, 2, 256, 128, generator=torch.Generator().manual_seed(s)).to(torch.float8_e5m2).to(DEV)
               for s in (1, 2, 3))
    O1, L1 = flash_attention_forward(Q, K, V, DEV, True, 0.1)
    O2, L2 = flash_attention_forward(Q, K, V, DEV, True, 0.1)
    assert torch.equal(O1.view(torch.uint8), O2.view(torch.uint8)) and torch.equal(L1, L2)
    with pytest.raises(TypeError):
        flash_attention_backward(Q, K, V, O1, O1, L1, DEV, False, True, 0.1)


# ------------------------------------------------------------------------------------------------ fused gather epilogue
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float8_e4m3fn])
def test_forward_writes_in_place_and_to_peer_windows(dtype):
    """fa_fwd_peers on one GPU: O goes to a strided window of a caller-owned 'gathered' buffer and to two more windows
    standing in for peer GPUs; all copies equal the plain forward bit for bit, L too."""
    B, H, N, D, Htot, h0 = 2, 3, 384, 128, 8, 2
    g = torch.Generator().manual_seed(5)
    Q, K, V = (torch.randn(B, H, N, D, generator=g).to(dtype).to(DEV) for _ in range(3))
    want_O, want_L = _native.forward(Q, K, V, True, 0.09)
    bufs = [torch.zeros(B, Htot, N, D, dtype=torch.uint8 if dtype.itemsize == 1 else dtype, device=DEV).view(dtype)
            for _ in range(3)]
    views = [b[:, h0:h0 + H] for b in bufs]
    none, L = _native.forward(Q, K, V, True, 0.09, out=(views[0].data_ptr(), tuple(views[0].stride())),
                              peer_ptrs=[v.data_ptr() for v in views[1:]])
    torch.cuda.synchronize()
    assert none is None and torch.equal(L, want_L)
    raw = lambda t: t.contiguous().view(torch.uint8)
    for b, v in zip(bufs, views):
        assert torch.equal(raw(v), raw(want_O))
        assert not raw(b[:, :h0]).any() and not raw(b[:, h0 + H:]).any()   # nothing outside the window


# ------------------------------------------------------------------------------------------------ the reference's own script
def test_reference_gradcheck_script_runs_unmodified(capsys):
    """src/test_torch.py of the reference, byte for byte (baseline/_ref/src, an untouched copy), with only sys.path
    pointing `flash_attention_torch` at flash_attention_dlrs_b200/compat: both gradchecks must report success."""
    import runpy
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = os.path.join(root, "baseline", "_ref", "src", "test_torch.py")
    if not os.path.exists(script):
        pytest.skip("baseline/_ref/src not present (tools/fetch_reference.sh)")
    compat = os.path.join(root, "flash_attention_dlrs_b200", "compat")
    sys.path.insert(0, compat)
    try:
        runpy.run_path(script, run_name="__main__")
    finally:
        sys.path.remove(compat)
    out = capsys.readouterr().out
    assert "Non-deterministic backwards test successful" in out, out
    assert "Deterministic backwards test successful" in out, out


def test_reference_correctness_script_runs_unmodified(capsys):
    """src/test_correctness.py of the reference, byte for byte (baseline/_ref/src): 200 seeds x (forward, dQ, dK, dV with
    deterministic=False, dQ, dK, dV with deterministic=True) against torch SDPA + autograd at the reference's own
    tolerances (test_correctness.py:28-76).  Only sys.path decides that `flash_attention_wrappers` is this package's
    compat shim; every one of the seven summary lines must read 200 out of 200."""
    import hashlib
    import runpy
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = os.path.join(root, "baseline", "_ref", "src", "test_correctness.py")
    if not os.path.exists(script):
        pytest.skip("baseline/_ref/src not present (tools/fetch_reference.sh)")
    assert hashlib.sha1(open(script, "rb").read()).hexdigest() == "2fbd97fea76fb687e152a580e04fa8359d08e34c"
    compat = os.path.join(root, "flash_attention_dlrs_b200", "compat")
    sys.path.insert(0, compat)
    for mod in ("flash_attention_wrappers", "flash_attention_torch"):   # a stale import must not decide the module
        sys.modules.pop(mod, None)
    try:
        runpy.run_path(script, run_name="__main__")
        import flash_attention_wrappers as w
        assert os.path.dirname(os.path.abspath(w.__file__)) == compat, w.__file__
    finally:
        sys.path.remove(compat)
    out = capsys.readouterr().out
    for what in ("forward", "Q backward", "K backward", "V backward", "Q deterministic backward",
                 "K deterministic backward", "V deterministic backward"):
        assert f"200 out of 200 {what} tests succeeded!" in out, out


def _free_port():
    """A TCP port nobody listens on right now (fixed ports collide when two test runs share a box)."""
    import socket
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        return sk.getsockname()[1]


def _run_peer_gather(nproc, port):
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc), "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(root, "tools", "multi_gpu_peer_gather.py"),
           "1", "8", "2048", "128"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert out.returncode == 0 and lines, out.stdout[-2000:] + out.stderr[-2000:]
    res = json.loads(lines[-1])
    flags = {k: v for k, v in res.items() if k.endswith("_equals_nccl_bitwise")}
    assert flags and all(flags.values()), res
    return res


def test_fused_gather_epilogue_single_rank():
    """PeerGatherBuffer end to end with one rank (torchrun, symmetric-memory rendezvous, kernel writing O in place at
    the buffer's address, barrier): equals the plain forward bit for bit."""
    _run_peer_gather(1, _free_port())


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs on one box")
def test_fused_gather_epilogue_matches_nccl_all_gather_two_gpus():
    """Two ranks (torchrun, NCCL for the rendezvous and the comparison path): the gathered O written by the forward
    kernels' epilogues (NVLS multicast and / or P2P stores) equals forward + NCCL all-gather bit for bit."""
    _run_peer_gather(2, _free_port())


# ------------------------------------------------------------------------------------------------ key-padding mask
def _varlen_reference(Q, K, V, dO, lens, scale, causal):
    """Ground truth for per-batch valid lengths: the oracle on each batch element's first lens[b] tokens, zeros beyond."""
    ref = {k: torch.zeros(Q.shape, dtype=torch.float64) for k in ("O", "dQ", "dK", "dV")}
    ref["L"] = torch.zeros(*Q.shape[:3], 1, dtype=torch.float64)
    for b, n in enumerate(lens):
        if n == 0:
            continue
        r = orc.attention_grads_fp64(Q[b:b + 1, :, :n].float(), K[b:b + 1, :, :n].float(), V[b:b + 1, :, :n].float(),
                                     dO[b:b + 1, :, :n].float(), scale, causal)
        for k in ref:
            ref[k][b:b + 1, :, :n] = r[k]
    return ref


@pytest.mark.parametrize("dtype,D", [(torch.bfloat16, 128), (torch.float16, 64), (torch.float32, 64)])
@pytest.mark.parametrize("causal", [False, True])
def test_key_padding_mask_per_batch_lengths(dtype, D, causal):
    """seqlens: every batch element has its own valid length; the padded positions hold arbitrary data (not zeros) and
    must influence nothing; O / L / gradients beyond the length are zero; backward stays bit-identical."""
    B, H, N = 5, 2, 384
    lens = [384, 200, 1, 129, 0]
    scale = 1.0 / math.sqrt(D)
    Q, K, V, dO = make_inputs(31, B, H, N, D, dtype)
    sl = torch.tensor(lens, dtype=torch.int32)
    q, k, v, do = (t.to(DEV) for t in (Q, K, V, dO))
    O, L = flash_attention_forward(q, k, v, DEV, causal, scale, sl)
    g = flash_attention_backward(q, k, v, O, do, L, DEV, True, causal, scale, sl)
    g2 = flash_attention_backward(q, k, v, O, do, L, DEV, True, causal, scale, sl)
    torch.cuda.synchronize()
    for a, b in zip(g, g2):
        assert torch.equal(a, b)
    ref = _varlen_reference(Q, K, V, dO, lens, scale, causal)
    tol_o = 1e-4 if dtype == torch.float32 else 2e-3
    for b, n in enumerate(lens):
        got_o, want_o = O[b, :, :n].cpu().double(), ref["O"][b, :, :n]
        if n:
            slack = 0
            if dtype != torch.float32:
                p_absv = orc.reference_sdpa(Q[b:b + 1, :, :n].float(), K[b:b + 1, :, :n].float(),
                                            V[b:b + 1, :, :n].float().abs(), scale, causal)[0].double()
                slack = out_half_ulp(want_o, dtype) + 2.0 ** -(MANT_BITS[dtype] + 2) * p_absv
            assert ((got_o - want_o).abs() <= tol_o + slack).all(), (b, n)
            assert (L[b, :, :n].cpu().double() - ref["L"][b, :, :n]).abs().max() <= tol_o, (b, n)
        assert not O[b, :, n:].float().abs().any() and not L[b, :, n:].abs().any(), (b, n)   # padded rows: zero
    for name, got in zip(("dQ", "dK", "dV"), g):
        assert rel_err(got.cpu(), ref[name]) <= (1e-4 if dtype == torch.float32 else 1e-2), name
        for b, n in enumerate(lens):
            assert not got[b, :, n:].float().abs().any(), (name, b, n)


def test_key_padding_mask_autograd_and_full_length_equivalence():
    """FlashAttention.apply(..., seqlens): gradients flow; seqlens = N everywhere reproduces the unmasked result bitwise."""
    B, H, N, D = 2, 2, 256, 128
    Q, K, V, dO = (t.to(DEV) for t in make_inputs(33, B, H, N, D, torch.bfloat16))
    q, k, v = (t.clone().requires_grad_(True) for t in (Q, K, V))
    O = FlashAttention.apply(q, k, v, True, 0.09, torch.tensor([N, N]))
    O.backward(dO)
    q2, k2, v2 = (t.clone().requires_grad_(True) for t in (Q, K, V))
    O2 = FlashAttention.apply(q2, k2, v2, True, 0.09)
    O2.backward(dO)
    assert torch.equal(O, O2)
    for a, b in ((q.grad, q2.grad), (k.grad, k2.grad), (v.grad, v2.grad)):
        assert torch.equal(a, b)
    with pytest.raises(ValueError):
        FlashAttention.apply(Q, K, V, True, 0.09, torch.tensor([N]))


# ------------------------------------------------------------------------------------------------ ring attention pieces
@pytest.mark.parametrize("dtype,D", [(torch.bfloat16, 128), (torch.float16, 64)])
@pytest.mark.parametrize("causal", [False, True])
def test_ring_attention_virtual_ranks_on_one_gpu(dtype, D, causal):
    """The arithmetic of the sequence-parallel path on one GPU: G virtual ranks run the schedule of ring.py in turn (no
    communication) with the CUDA kernels — per-shard fa_fwd / fa_bwd, fa_merge_partial, fa_accumulate, fa_round_rows —
    and the assembled O, L, dQ, dK, dV must match full attention over the whole sequence."""
    from flash_attention_dlrs_b200.ring import CudaOps as ops
    G, B, H, n = 4, 1, 2, 256
    N, scale = G * n, 1.0 / math.sqrt(D)
    Q, K, V, dO = make_inputs(41, B, H, N, D, dtype)
    q, k, v, do = (t.to(DEV) for t in (Q, K, V, dO))
    sh = lambda t, r: t[:, :, r * n:(r + 1) * n].contiguous()
    O = torch.empty_like(q)
    L = torch.empty(B, H, N, device=DEV)
    for r in range(G):
        o_acc = torch.empty(B, H, n, D, device=DEV)
        l_acc = torch.empty(B, H, n, device=DEV)
        first = True
        for s in range(G):
            c = (r - s) % G
            if causal and c > r:
                continue
            o_p, l_p = ops.fwd(sh(q, r), sh(k, c), sh(v, c), causal and c == r, scale)
            ops.merge(o_acc, l_acc, o_p, l_p, first)
            first = False
        O[:, :, r * n:(r + 1) * n] = ops.round(o_acc, dtype)
        L[:, :, r * n:(r + 1) * n] = l_acc
    dq = torch.empty(B, H, N, D, device=DEV)
    dk = torch.zeros(B, H, N, D, device=DEV)
    dv = torch.zeros(B, H, N, D, device=DEV)
    for r in range(G):
        delta = ops.delta(sh(O, r), sh(do, r))
        dq_acc = torch.empty(B, H, n, D, device=DEV)
        first = True
        for s in range(G):
            c = (r - s) % G
            if causal and c > r:
                continue
            g = ops.bwd(sh(q, r), sh(k, c), sh(v, c), sh(O, r), sh(do, r), sh(L, r), causal and c == r, scale, delta)
            ops.accumulate(dq_acc, g[0], first)
            first = False
            for acc, part in ((dk, g[1]), (dv, g[2])):     # the travelling accumulators, here simply addressed by shard
                tmp = acc[:, :, c * n:(c + 1) * n].contiguous()
                ops.accumulate(tmp, part, False)
                acc[:, :, c * n:(c + 1) * n] = tmp
        dq[:, :, r * n:(r + 1) * n] = dq_acc
    torch.cuda.synchronize()
    ref = orc.attention_grads_fp64(Q.float(), K.float(), V.float(), dO.float(), scale, causal)
    bound = o_bound(Q, K, V, ref["O"], scale, causal, dtype) + 2.0 ** -(MANT_BITS[dtype] + 1) * \
        orc.reference_sdpa(Q.float(), K.float(), V.float().abs(), scale, causal).double()   # + rounding of the partials
    assert ((O.cpu().double() - ref["O"]).abs() <= bound).all()
    assert (L.cpu().double().unsqueeze(-1) - ref["L"]).abs().max() <= 2e-3
    for name, got in (("dQ", dq), ("dK", dk), ("dV", dv)):
        assert rel_err(got.cpu(), ref[name]) <= 1e-2, name


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs on one box")
def test_ring_attention_two_gpus():
    """Two ranks over NCCL: ring forward + backward against full attention computed locally on every rank."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), os.path.join(root, "tools", "multi_gpu_ring.py"), "1", "4", "2048", "128"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert out.returncode == 0 and lines, out.stdout[-2000:] + out.stderr[-2000:]
    for line in lines:
        res = json.loads(line)
        assert res["ok"], res


# ------------------------------------------------------------------------------------------------ extremes
@pytest.mark.parametrize("shape", [(0, 4, 128, 64), (2, 0, 128, 64), (2, 4, 0, 64)])
def test_empty_inputs_return_empty_outputs(shape):
    Q = torch.empty(shape, dtype=torch.bfloat16, device=DEV)
    O, L = flash_attention_forward(Q, Q, Q, DEV, True, 0.1)
    assert O.shape == Q.shape and L.shape == (*shape[:3], 1)
    g = flash_attention_backward(Q, Q, Q, O, Q, L, DEV, False, True, 0.1)
    assert all(t.shape == Q.shape for t in g)


def test_long_ragged_sequence_against_an_independent_gpu_implementation():
    """N = 65536 + 77 (ragged, 513 key blocks), too large for the fp64 oracle: cross-check against torch's own flash
    SDPA on the GPU (an independent implementation; looser tolerance than the oracle tests), and bit-identical reruns."""
    B, H, N, D = 1, 2, 65536 + 77, 128
    scale = D ** -0.5
    g = torch.Generator(device=DEV).manual_seed(5)
    Q, K, V, dO = (torch.randn(B, H, N, D, generator=g, device=DEV, dtype=torch.float32).to(torch.bfloat16) for _ in range(4))
    O, L = flash_attention_forward(Q, K, V, DEV, True, scale)
    grads = flash_attention_backward(Q, K, V, O, dO, L, DEV, False, True, scale)
    again = flash_attention_backward(Q, K, V, O, dO, L, DEV, False, True, scale)
    for a, b in zip(grads, again):
        assert torch.equal(a, b)
    q, k, v = (t.clone().requires_grad_(True) for t in (Q, K, V))
    ref = torch.nn.functional.scaled_dot_product_attention(q, k, v, is_causal=True, scale=scale)
    rg = torch.autograd.grad(ref, (q, k, v), dO)
    assert (O.float() - ref.float()).abs().max().item() <= 3e-2
    for name, a, b in zip(("dQ", "dK", "dV"), grads, rg):
        assert rel_err(a.cpu(), b.detach().cpu()) <= 2e-2, name
    assert torch.isfinite(L).all()


def test_bnhd_storage_viewed_as_bhnd_needs_no_copy_and_matches():
    """(B, N, H, D) storage (the layout flash_attn uses) viewed as (B, H, N, D): the strides go straight into the TMA
    tensor maps; results equal the contiguous run bit for bit, forward and backward."""
    B, H, N, D = 2, 4, 384, 128
    Q, K, V, dO = (t.to(DEV) for t in make_inputs(51, B, H, N, D, torch.bfloat16))
    tq, tk, tv, tdo = (t.transpose(1, 2).contiguous().transpose(1, 2) for t in (Q, K, V, dO))   # BNHD storage
    assert not tq.is_contiguous() and _native._kernel_ready(tq).data_ptr() == tq.data_ptr()
    O1, L1 = flash_attention_forward(Q, K, V, DEV, True, 0.09)
    O2, L2 = flash_attention_forward(tq, tk, tv, DEV, True, 0.09)
    assert torch.equal(O1, O2) and torch.equal(L1, L2)
    g1 = flash_attention_backward(Q, K, V, O1, dO, L1, DEV, False, True, 0.09)
    g2 = flash_attention_backward(tq, tk, tv, O2, tdo, L2, DEV, False, True, 0.09)
    for a, b in zip(g1, g2):
        assert torch.equal(a, b)


# ------------------------------------------------------------------------------------------------ in-kernel dropout
def _run_dropout(Q, K, V, dO, causal, scale, p, seed, seqlens=None, mask=None):
    q, k, v, do = (t.to(DEV) for t in (Q, K, V, dO))
    mask = None if mask is None else mask.to(DEV)
    O, L = flash_attention_forward(q, k, v, DEV, causal, scale, seqlens, p, seed, mask)
    g = flash_attention_backward(q, k, v, O, do, L, DEV, True, causal, scale, seqlens, p, seed, mask)
    torch.cuda.synchronize()
    return (O.cpu(), L.cpu()) + tuple(t.cpu() for t in g)


@pytest.mark.parametrize("dtype,D", [(torch.bfloat16, 128), (torch.float32, 128)])
@pytest.mark.parametrize("causal", [False, True])
def test_dropout_mask_is_bit_exact_in_all_three_kernels(dtype, D, causal):
    """The keep mask every kernel regenerates equals the oracle's, entry by entry.  Q = 0 makes P uniform, and
    identity-like V / dO / K make single entries of the dropped P visible: O[i, d] = keep(i, d) P rp (forward, row walk),
    dV[j, d] = keep(d, j) P rp (dK/dV kernel, column walk), sign(dQ[i, d]) = +/- for kept / dropped (dQ kernel)."""
    B, H, N, p, seed = 1, 2, 256, 0.25, 987654321012345
    keep = orc.dropout_keep_mask(seed, B, H, N, p)                      # (B,H,N,N)
    eye = torch.eye(128)
    Q = torch.zeros(B, H, N, D, dtype=dtype)
    ones = torch.ones(B, H, N, D, dtype=dtype)
    tri = torch.ones(N, N, dtype=torch.bool).tril() if causal else torch.ones(N, N, dtype=torch.bool)
    for sel in (0, 1):
        half = torch.zeros(N, D)
        half[sel * 128:(sel + 1) * 128] = eye
        half = half.to(dtype).expand(B, H, N, D).contiguous()
        ks = slice(sel * 128, (sel + 1) * 128)
        # forward: V = identity on keys [ks] -> O[i, d] = keep(i, ks[d]) * P(i) * rp
        # (a causally masked entry whose exponential went through the FMA-pipe polynomial is 2^-125, not 0: "> 1e-30")
        O = _run_dropout(Q, ones, half, ones, causal, 1.0, p, seed)[0]
        assert torch.equal(O.float() > 1e-30, keep[..., ks] & tri[:, ks]), ("forward", sel)
        # dK/dV kernel: dO = identity on queries [ks] -> dV[j, d] = keep(ks[d], j) * P * rp
        dV = _run_dropout(Q, ones, ones, half, causal, 1.0, p, seed)[4]
        assert torch.equal(dV != 0, (keep[:, :, ks, :] & tri[ks, :]).transpose(-1, -2)), ("dkdv", sel)
        # dQ kernel: K = identity on keys [ks], V = 1, dO = 1/D -> dP = 1, delta ~ 1, dQ[i, d] = P (keep rp - delta)
        dQ = _run_dropout(Q, half, ones, ones / D, causal, 1.0, p, seed)[2]
        # (rows that see few keys can have every key kept, i.e. delta = rp and dS = 0: start the causal check at row 64)
        valid = tri[:, ks].clone()
        valid[:64 if causal else 0] = False
        valid = valid.expand(B, H, N, 128)
        assert torch.equal((dQ > 0)[valid], keep[..., ks][valid]), ("dq", sel)
        assert not (dQ == 0)[valid].any() and not dQ[~tri[:, ks].expand(B, H, N, 128)].any()


@pytest.mark.parametrize("dtype,D", [(torch.bfloat16, 128), (torch.float16, 64), (torch.bfloat16, 64), (torch.float32, 64)])
@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("N", [384, 200])
def test_dropout_parity(dtype, D, causal, N):
    """O, L, dQ, dK, dV with in-kernel dropout against the float64 closed form with the oracle's mask."""
    B, H, p, seed = 2, 3, 0.2, 20240 + N
    scale = 1.0 / math.sqrt(D)
    Q, K, V, dO = make_inputs(77, B, H, N, D, dtype)
    O, L, dQ, dK, dV = _run_dropout(Q, K, V, dO, causal, scale, p, seed)
    keep = orc.dropout_keep_mask(seed, B, H, N, p)
    ref = orc.attention_dropout_grads_fp64(Q.float(), K.float(), V.float(), dO.float(), scale, causal, keep, p)
    rp = 256.0 / (256.0 - orc.dropout_threshold(p))
    o_err = (O.double() - ref["O"]).abs()
    if dtype == torch.float32:
        assert o_err.max().item() <= 1e-4 and (L.double() - ref["L"]).abs().max().item() <= 1e-4
    else:
        p_absv = orc.reference_sdpa(Q.float(), K.float(), V.float().abs(), scale, causal).double()
        bound = rp * (2e-3 + 2.0 ** -(MANT_BITS[dtype] + 2) * p_absv) + out_half_ulp(ref["O"], dtype)
        assert (o_err <= bound).all(), f"O err {o_err.max().item():.3e}"
        assert (L.double() - ref["L"]).abs().max().item() <= 2e-3
    for name, got in (("dQ", dQ), ("dK", dK), ("dV", dV)):
        e = rel_err(got, ref[name])
        assert e <= (2e-4 if dtype == torch.float32 else 1e-2), f"{name} rel err {e:.3e}"
    # L is the logsumexp of the undropped scores: bitwise the same as without dropout
    q, k, v = (t.to(DEV) for t in (Q, K, V))
    assert torch.equal(flash_attention_forward(q, k, v, DEV, causal, scale)[1].cpu(), L)


def test_dropout_with_key_padding_and_padded_head_size():
    """dropout composes with seqlens (mask indexed by absolute positions) and with a head size that needs padding."""
    B, H, N, d, p, seed = 3, 2, 300, 40, 0.3, 5150
    lens = [300, 130, 77]
    scale = 1.0 / math.sqrt(d)
    Q, K, V, dO = make_inputs(78, B, H, N, d, torch.bfloat16)
    O, L, dQ, dK, dV = _run_dropout(Q, K, V, dO, True, scale, p, seed, torch.tensor(lens, dtype=torch.int32))
    keep = orc.dropout_keep_mask(seed, B, H, N, p)
    for b, n in enumerate(lens):
        r = orc.attention_dropout_grads_fp64(Q[b:b + 1, :, :n].float(), K[b:b + 1, :, :n].float(), V[b:b + 1, :, :n].float(),
                                             dO[b:b + 1, :, :n].float(), scale, True, keep[b:b + 1, :, :n, :n], p)
        assert (O[b:b + 1, :, :n].double() - r["O"]).abs().max() <= 3e-2   # bf16 output rounding at |O| ~ 3
        for name, got in (("dQ", dQ), ("dK", dK), ("dV", dV)):
            assert rel_err(got[b:b + 1, :, :n], r[name]) <= 1e-2, (name, b)
            assert not got[b, :, n:].float().abs().any()


def test_dropout_autograd_seeding_and_determinism():
    B, H, N, D = 2, 2, 512, 128
    Q, K, V, dO = (t.to(DEV) for t in make_inputs(79, B, H, N, D, torch.bfloat16))

    def run(*extra):
        q, k, v = (t.clone().requires_grad_(True) for t in (Q, K, V))
        O = FlashAttention.apply(q, k, v, True, 0.09, None, *extra)
        O.backward(dO)
        return O.detach(), q.grad, k.grad, v.grad

    a, b, c = run(0.1, 11), run(0.1, 11), run(0.1, 12)
    for x, y in zip(a, b):
        assert torch.equal(x, y)                       # same seed: bit-identical forward and backward
    assert not torch.equal(a[0], c[0])                 # another seed: another mask
    for x, y in zip(run(0.0, 11), run()):
        assert torch.equal(x, y)                       # p = 0 is the no-dropout kernel
    torch.manual_seed(123)
    d1 = run(0.1)
    torch.manual_seed(123)
    d2 = run(0.1)
    assert torch.equal(d1[0], d2[0]) and torch.equal(d1[1], d2[1])   # default seed follows torch.manual_seed
    assert not torch.equal(run(0.1)[0], d1[0])                       # and advances from call to call
    # unbiased: the mean over masks approaches the undropped output
    base = run()[0].float()
    acc = torch.zeros_like(base)
    for s in range(64):
        acc += FlashAttention.apply(Q, K, V, True, 0.09, None, 0.5, 1000 + s).float()
    late = slice(256, None)   # rows with many keys: the mask noise averages out
    assert (acc[:, :, late] / 64 - base[:, :, late]).abs().mean().item() < 0.03
    with pytest.raises(TypeError):
        _native.forward(Q.to(torch.float8_e4m3fn), K.to(torch.float8_e4m3fn), V.to(torch.float8_e4m3fn), True, 0.09,
                        dropout_p=0.5, dropout_seed=1)
    with pytest.raises(_lib.FlashAttentionLibraryError):
        _native.backward(Q, K, V, a[0], dO, torch.zeros(B, H, N, device=DEV), True, 0.09, which=_native.BWD_FUSED,
                         dropout_p=0.5, dropout_seed=1)


# ------------------------------------------------------------------------------------------------ randomized combinations
@pytest.mark.parametrize("case", range(40))
def test_randomized_feature_combinations(case):
    """Seeded random draws over everything the default path takes at once — dtype, head size (padded or not), ragged N,
    causal, scale, per-batch lengths, dropout, an arbitrary mask — against the float64 closed form evaluated per batch
    element."""
    rng = np.random.default_rng(1000 + case)
    dtype = (torch.bfloat16, torch.float16, torch.float32)[case % 3]
    d = int(rng.choice([16, 24, 40, 64, 72, 128]))
    N = int(rng.choice([1, 17, 127, 128, 129, 255, 300, 511, 640]))
    B, H = int(rng.integers(1, 4)), int(rng.integers(1, 4))
    causal = bool(rng.integers(0, 2))
    scale = float(rng.choice([1.0 / math.sqrt(d), 0.05, 0.3]))
    p = float(rng.choice([0.0, 0.0, 0.1, 0.5]))
    seed = int(rng.integers(0, 2 ** 62))
    lens = [int(x) for x in rng.integers(0, N + 1, size=B)] if rng.integers(0, 2) else None
    mask = None
    if rng.integers(0, 2):
        shape = [(N, N), (B, N, N), (1, H, N, N), (B, H, N, N)][int(rng.integers(0, 4))]
        mask = torch.from_numpy(rng.random(shape) < 0.75)
        mask4 = mask.reshape((1,) * (4 - mask.dim()) + tuple(mask.shape)) if mask.dim() != 3 else mask[:, None]
    Q, K, V, dO = make_inputs(2000 + case, B, H, N, d, dtype)
    sl = None if lens is None else torch.tensor(lens, dtype=torch.int32)
    O, L, dQ, dK, dV = _run_dropout(Q, K, V, dO, causal, scale, p, seed, sl, mask)
    keep = orc.dropout_keep_mask(seed, B, H, N, p)
    rp = 256.0 / (256.0 - orc.dropout_threshold(p))
    what = (f"case {case}: {dtype} d={d} N={N} B={B} H={H} causal={causal} scale={scale:.3f} p={p} lens={lens} "
            f"mask={None if mask is None else tuple(mask.shape)}")
    ref = {k: torch.zeros(Q.shape, dtype=torch.float64) for k in ("O", "dQ", "dK", "dV")}
    ref["L"] = torch.zeros(B, H, N, 1, dtype=torch.float64)
    bound = torch.zeros(Q.shape, dtype=torch.float64)
    for b in range(B):
        n = N if lens is None else lens[b]
        if n == 0:
            continue
        qb, kb, vb, dob = (t[b:b + 1, :, :n].float() for t in (Q, K, V, dO))
        mb = None if mask is None else mask4[(b if mask4.shape[0] > 1 else 0):(b if mask4.shape[0] > 1 else 0) + 1, :, :n, :n]
        r = orc.attention_dropout_grads_fp64(qb, kb, vb, dob, scale, causal, keep[b:b + 1, :, :n, :n], p, mb)
        for k in ref:
            ref[k][b:b + 1, :, :n] = torch.nan_to_num(r[k], neginf=0.0) if k == "L" else r[k]
        if dtype != torch.float32:
            vis = torch.ones(n, n, dtype=torch.bool).tril() if causal else torch.ones(n, n, dtype=torch.bool)
            vis = vis if mb is None else (vis & mb)
            Pm = torch.softmax((scale * qb.double() @ kb.double().transpose(-1, -2)).masked_fill(~vis, -math.inf), -1)
            p_absv = torch.nan_to_num(Pm, nan=0.0) @ vb.double().abs()
            # worst-case rounding of P to the input dtype (relative 2^-(mant+1)): with the sharp softmaxes drawn here one
            # key can carry the whole row, so the rounding errors do not average out as in the fixed-scale tests above
            bound[b:b + 1, :, :n] = rp * (2e-3 + 2.0 ** -(MANT_BITS[dtype] + 1) * p_absv) + out_half_ulp(r["O"], dtype)
        else:
            bound[b:b + 1, :, :n] = 1e-4 * rp
    assert ((O.double() - ref["O"]).abs() <= bound).all(), what + f" O err {(O.double() - ref['O']).abs().max():.3e}"
    L = torch.nan_to_num(L, neginf=0.0)   # queries with no visible key: L = -inf on both sides
    assert (L.double() - ref["L"]).abs().max() <= (1e-4 if dtype == torch.float32 else 2e-3), what
    for name, got in (("dQ", dQ), ("dK", dK), ("dV", dV)):
        if ref[name].abs().max() < 1e-9:
            # one visible key per row: dS = P o (dP - delta) is identically zero in exact arithmetic, what the kernels
            # return is the rounding of O inside delta -> absolute check
            e = (got.double() - ref[name]).abs().max().item()
            assert e <= (1e-4 if dtype == torch.float32 else 2e-2), what + f" {name} abs err {e:.3e} (zero reference)"
            continue
        e = rel_err(got, ref[name])
        assert e <= (3e-4 if dtype == torch.float32 else 1e-2), what + f" {name} rel err {e:.3e}"


# ------------------------------------------------------------------------------------------------ arbitrary attention mask
def _run_masked(Q, K, V, dO, causal, scale, mask, seqlens=None, p=0.0, seed=None):
    q, k, v, do = (t.to(DEV) for t in (Q, K, V, dO))
    from flash_attention_dlrs_b200 import AttentionMask
    am = AttentionMask(mask.to(DEV))
    O, L = flash_attention_forward(q, k, v, DEV, causal, scale, seqlens, p, seed, am)
    g = flash_attention_backward(q, k, v, O, do, L, DEV, True, causal, scale, seqlens, p, seed, am)
    g2 = flash_attention_backward(q, k, v, O, do, L, DEV, True, causal, scale, seqlens, p, seed, am)
    torch.cuda.synchronize()
    for a, b in zip(g, g2):
        assert torch.equal(a, b)   # deterministic with a mask as well
    return (O.cpu(), L.cpu()) + tuple(t.cpu() for t in g)


def _check_masked(Q, K, V, dO, causal, scale, mask, got, dtype, p=0.0, seed=0):
    B, H, N, _ = Q.shape
    O, L, dQ, dK, dV = got
    keep = orc.dropout_keep_mask(seed, B, H, N, p)
    ref = orc.attention_dropout_grads_fp64(Q.float(), K.float(), V.float(), dO.float(), scale, causal, keep, p, mask)
    full = mask.expand(B, H, N, N) & (torch.ones(N, N, dtype=torch.bool).tril() if causal else True)
    empty = ~full.any(-1)                                              # queries with no visible key
    assert not O[empty].float().abs().any() and (L.squeeze(-1)[empty] == -math.inf).all()
    o_err = (O.double() - ref["O"]).abs()
    seen = ~empty
    if dtype == torch.float32:
        assert o_err.max() <= 1e-4 and (L.squeeze(-1)[seen].double() - ref["L"].squeeze(-1)[seen]).abs().max() <= 1e-4
    else:
        # P|V| over the visible keys only (an upper bound: the full softmax with |V| is not, the mask renormalises)
        Pm = torch.softmax((scale * Q.double() @ K.double().transpose(-1, -2)).masked_fill(~full, -math.inf), -1)
        p_absv = torch.nan_to_num(Pm, nan=0.0) @ V.double().abs()
        rp = 256.0 / (256.0 - orc.dropout_threshold(p))
        bound = rp * (2e-3 + 2.0 ** -(MANT_BITS[dtype] + (2 if p == 0.0 else 1)) * p_absv) + out_half_ulp(ref["O"], dtype)
        assert (o_err <= bound).all(), f"O err {o_err.max().item():.3e}"
        assert (L.squeeze(-1)[seen].double() - ref["L"].squeeze(-1)[seen]).abs().max() <= 2e-3
    for name, g in (("dQ", dQ), ("dK", dK), ("dV", dV)):
        assert torch.isfinite(g.float()).all(), name
        e = rel_err(g, ref[name])
        assert e <= (3e-4 if dtype == torch.float32 else 1e-2), f"{name} rel err {e:.3e}"


@pytest.mark.parametrize("dtype,D", [(torch.bfloat16, 128), (torch.float16, 64), (torch.float32, 64)])
@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("kind", ["random", "blocks", "band"])
def test_attention_mask_parity(dtype, D, causal, kind):
    """Arbitrary bool masks (True = attend) ANDed with the causal flag: random 70 % dense, block-sparse with whole
    128 x 128 blocks (and whole query rows) masked out, and a sliding-window band; ragged N; one mask for all heads."""
    B, H, N = 2, 3, 392
    g = torch.Generator().manual_seed(5)
    if kind == "random":
        mask = torch.rand(B, 1, N, N, generator=g) < 0.7
        mask[0, 0, 100] = False                                   # a query that sees nothing
        mask[1, 0, :, 200] = False                                # a key nobody attends to
    elif kind == "blocks":
        blk = torch.rand(1, H, 4, 4, generator=g) < 0.5
        blk[0, :, 2, :] = False                                   # query block 2 of every head is empty
        blk[0, :, 0, 0] = True
        mask = blk.repeat_interleave(128, -1).repeat_interleave(128, -2)[..., :N, :N]
    else:
        i = torch.arange(N)
        mask = ((i[:, None] - i[None, :]).abs() <= 37)[None, None]   # (1,1,N,N) sliding window
    scale = 1.0 / math.sqrt(D)
    Q, K, V, dO = make_inputs(91, B, H, N, D, dtype)
    got = _run_masked(Q, K, V, dO, causal, scale, mask)
    _check_masked(Q, K, V, dO, causal, scale, mask, got, dtype)


def test_attention_mask_equivalences_and_autograd():
    """A tril mask reproduces the causal kernel, an all-True mask the unmasked one (same values up to the exponent
    polynomial the unmasked forward mixes in); masks given as (N,N) / (B,N,N) / AttentionMask; gradients flow."""
    B, H, N, D = 2, 2, 300, 128
    Q, K, V, dO = (t.to(DEV) for t in make_inputs(92, B, H, N, D, torch.bfloat16))

    def run(causal, mask):
        q, k, v = (t.clone().requires_grad_(True) for t in (Q, K, V))
        O = FlashAttention.apply(q, k, v, causal, 0.09, None, 0.0, None, mask)
        O.backward(dO)
        return [O.detach().float(), q.grad.float(), k.grad.float(), v.grad.float()]

    tril = torch.ones(N, N, dtype=torch.bool, device=DEV).tril()
    for a, b in zip(run(True, None), run(False, tril)):
        assert (a - b).abs().max() <= 2e-2 * b.abs().max()
    for a, b in zip(run(False, None), run(False, torch.ones(B, N, N, dtype=torch.bool, device=DEV))):
        assert (a - b).abs().max() <= 2e-2 * b.abs().max()
    from flash_attention_dlrs_b200 import AttentionMask
    am = AttentionMask(tril)
    for a, b in zip(run(False, tril), run(True, am)):     # causal AND tril == tril; packed mask reused
        assert torch.equal(a, b)
    with pytest.raises(ValueError):
        FlashAttention.apply(Q, K, V, False, 0.09, None, 0.0, None, torch.ones(N + 1, N + 1, dtype=torch.bool, device=DEV))
    with pytest.raises(ValueError):
        FlashAttention.apply(Q, K, V, False, 0.09, None, 0.0, None, torch.ones(3, N, N, dtype=torch.bool, device=DEV))


def test_attention_mask_with_seqlens_and_dropout():
    """mask AND key padding; mask AND dropout (float32 and 16-bit kernels)."""
    B, H, N, D = 3, 2, 260, 64
    g = torch.Generator().manual_seed(6)
    mask = torch.rand(B, H, N, N, generator=g) < 0.6
    lens = [260, 131, 40]
    Q, K, V, dO = make_inputs(93, B, H, N, D, torch.float16)
    O, L, dQ, dK, dV = _run_masked(Q, K, V, dO, False, 0.125, mask, torch.tensor(lens, dtype=torch.int32))
    for b, n in enumerate(lens):
        sub = tuple(t[b:b + 1, :, :n] for t in (Q, K, V, dO))
        _check_masked(*sub, False, 0.125, mask[b:b + 1, :, :n, :n],
                      tuple(t[b:b + 1, :, :n] for t in (O, L, dQ, dK, dV)), torch.float16)
        for t in (O, dQ, dK, dV):
            assert not t[b, :, n:].float().abs().any()
    Q, K, V, dO = make_inputs(94, 2, 2, 200, 32, torch.float32)
    mask = torch.rand(2, 1, 200, 200, generator=g) < 0.5
    got = _run_masked(Q, K, V, dO, True, 0.2, mask, None, 0.3, 777)
    _check_masked(Q, K, V, dO, True, 0.2, mask, got, torch.float32, 0.3, 777)
    for dtype, D in ((torch.bfloat16, 128), (torch.float16, 64)):
        Q, K, V, dO = make_inputs(96, 2, 2, 300, D, dtype)
        mask = torch.rand(1, 2, 300, 300, generator=g) < 0.6
        mask[0, :, 128:256, :128] = False      # an empty block and
        mask[0, :, 128:256, 128:256] = True    # a fully visible one
        got = _run_masked(Q, K, V, dO, False, 1.0 / math.sqrt(D), mask, None, 0.25, 4711)
        _check_masked(Q, K, V, dO, False, 1.0 / math.sqrt(D), mask, got, dtype, 0.25, 4711)


@pytest.mark.parametrize("dtype,D", [(torch.bfloat16, 128), (torch.float16, 64), (torch.float32, 32)])
@pytest.mark.parametrize("causal", [False, True])
def test_attention_mask_block_skipping_long(dtype, D, causal):
    """Sparse masks over 11 x 11 blocks: most 128 x 128 blocks are empty and skipped (forward tiles whose two halves see
    different block lists, a leading run of skipped blocks, query / key blocks with nothing visible at all), so the
    skip paths of every pipeline role run through several ring wrap-arounds."""
    B, H, N = 1, 2, 1300
    i = torch.arange(N)
    band = (i[:, None] - i[None, :]).abs() <= 150                       # sliding window
    band[300:420] = False                                               # queries that see nothing (most of block 2, 3)
    band[:, 700:900] = False                                            # keys nobody sees (all of block 6)
    band[1000:1100, 0:50] = True                                        # a far-away "sink" block for late queries
    glob = torch.zeros(N, N, dtype=torch.bool)
    glob[640:768, :] = True                                             # one query block attends everywhere
    glob[:, 700:900] = False
    mask = torch.stack([band, band | glob])[None]                       # (1, H, N, N): a different mask per head
    scale = 1.0 / math.sqrt(D)
    Q, K, V, dO = make_inputs(95, B, H, N, D, dtype)
    got = _run_masked(Q, K, V, dO, causal, scale, mask)
    _check_masked(Q, K, V, dO, causal, scale, mask, got, dtype)
    # the block summary must not change any value: same run with skipping disabled
    from flash_attention_dlrs_b200 import AttentionMask
    am = AttentionMask(mask.to(DEV))
    am.blocks.fill_(1)
    q, k, v, do = (t.to(DEV) for t in (Q, K, V, dO))
    O2, L2 = flash_attention_forward(q, k, v, DEV, causal, scale, None, 0.0, None, am)
    g2 = flash_attention_backward(q, k, v, O2, do, L2, DEV, True, causal, scale, None, 0.0, None, am)
    for a, b in zip(got, (O2, L2, *g2)):
        assert torch.equal(a, b.cpu())


def test_attention_mask_at_the_512_block_limit_of_the_skip_lists():
    """N = 65536 = 512 key blocks, the capacity of the kernels' shared-memory block lists: a +-200 sliding window (most of
    the 512 x 512 blocks skipped, list indices up to 511).  Checked on windows of the sequence against the oracle on the
    corresponding sub-problem: interior queries see exactly the same keys there, so O / L / dQ (and dK / dV further
    inside) must agree."""
    from flash_attention_dlrs_b200 import AttentionMask
    B, H, N, D, W = 1, 1, 65536, 64, 200
    scale = 0.125
    g = torch.Generator().manual_seed(97)
    Q, K, V, dO = ((torch.randn(B, H, N, D, generator=g)).to(torch.bfloat16) for _ in range(4))
    am = AttentionMask.sliding_window(N, W, W, device=DEV)
    assert am.blocks.shape[-1] == 512 and (am.blocks > 0).float().mean() < 0.01
    q, k, v, do = (t.to(DEV) for t in (Q, K, V, dO))
    O, L = flash_attention_forward(q, k, v, DEV, False, scale, None, 0.0, None, am)
    dQ, dK, dV = flash_attention_backward(q, k, v, O, do, L, DEV, True, False, scale, None, 0.0, None, am)
    torch.cuda.synchronize()
    del am
    for s0 in (0, 32768 - 300, N - 1100):
        e0 = s0 + 1100
        sub = tuple(t[:, :, s0:e0] for t in (Q, K, V, dO))
        i = torch.arange(e0 - s0)
        m = ((i[:, None] - i[None, :]).abs() <= W)[None, None]
        keep = torch.ones(1, 1, e0 - s0, e0 - s0, dtype=torch.bool)
        ref = orc.attention_dropout_grads_fp64(*(t.float() for t in sub), scale, False, keep, 0.0, m)
        lo_q = 0 if s0 == 0 else W            # queries whose whole window lies inside the slice
        hi_q = (e0 - s0) if e0 == N else (e0 - s0 - W)
        lo_k = 0 if s0 == 0 else 2 * W        # keys whose queries are all such queries
        hi_k = (e0 - s0) if e0 == N else (e0 - s0 - 2 * W)
        got_o = O[:, :, s0:e0].cpu().double()
        assert (got_o[:, :, lo_q:hi_q] - ref["O"][:, :, lo_q:hi_q]).abs().max() <= 2e-2
        assert (L[:, :, s0:e0].cpu().double()[:, :, lo_q:hi_q] - ref["L"][:, :, lo_q:hi_q]).abs().max() <= 2e-3
        assert rel_err(dQ[:, :, s0:e0].cpu()[:, :, lo_q:hi_q], ref["dQ"][:, :, lo_q:hi_q]) <= 1e-2
        assert rel_err(dK[:, :, s0:e0].cpu()[:, :, lo_k:hi_k], ref["dK"][:, :, lo_k:hi_k]) <= 1e-2
        assert rel_err(dV[:, :, s0:e0].cpu()[:, :, lo_k:hi_k], ref["dV"][:, :, lo_k:hi_k]) <= 1e-2


@pytest.mark.parametrize("dtype,D", [(torch.bfloat16, 128), (torch.float16, 64), (torch.float32, 32)])
@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("left,right", [(100, 30), (0, 0), (5, 400), (1000, 1000)])
def test_band_mask_equals_the_same_mask_given_as_bytes(dtype, D, causal, left, right):
    """AttentionMask.sliding_window (no mask bytes: the kernels compute the band, the block summary is analytic) gives
    bit for bit what the same band gives as a byte mask, forward and backward, with dropout as well."""
    from flash_attention_dlrs_b200 import AttentionMask
    B, H, N = 2, 2, 700
    scale = 1.0 / math.sqrt(D)
    Q, K, V, dO = (t.to(DEV) for t in make_inputs(98, B, H, N, D, dtype))
    i = torch.arange(N, device=DEV)
    d = i[None, :] - i[:, None]
    dense = AttentionMask((d >= -left) & (d <= right))
    band = AttentionMask.sliding_window(N, left, right, device=DEV)
    for p_drop, seed in ((0.0, None), (0.2, 31337)):
        outs = []
        for am in (dense, band):
            O, L = flash_attention_forward(Q, K, V, DEV, causal, scale, None, p_drop, seed, am)
            g = flash_attention_backward(Q, K, V, O, dO, L, DEV, True, causal, scale, None, p_drop, seed, am)
            outs.append((O, L, *g))
        for a, b in zip(*outs):
            assert torch.equal(a, b)
