#!/usr/bin/env python
"""Sequence-length sweep and the other BASELINE.json configs, next to every baseline the reference's own bench lines up
(src/bench.py:38-42,67-89) and the ones this box adds:

    b200-cuda          this library through FlashAttention.apply (autograd and launch overhead included)
    reference-triton   the reference's own kernels, unmodified sources (fp16 / fp32, scale 1, non-causal only)
    openai-tutorial    the vendored tutorial kernel (flash_attention_openai_tutorial.py:438-520; fp16, causal flag)
    daolab-fa2         flash_attn 2.8 (sm_80 mma.sync kernels)
    torch-fa           torch SDPA, FLASH_ATTENTION backend (sm_80 kernels as well)
    torch-cudnn        torch SDPA, CUDNN_ATTENTION backend: cuDNN 9's Blackwell-native fused attention — the one
                       sm_100-native competitor in the image
    torch-xformers     torch SDPA, EFFICIENT_ATTENTION backend          (src/bench.py:78-79)
    torch-math         torch SDPA, MATH backend (N <= 4096: it materialises the N x N matrix; src/bench.py:80-81)
    cpu-torch          the reference's CPU ground-truth path (torch SDPA fp32 + autograd on the host cores) on a 2-head
                       subset, time scaled by B*H/2 (BASELINE.json configs[4] asks for it; stated as scaled)

Not part of the driver contract (that is bench.py); writes

    bench_out/fused-attention-B{B}-H{H}-d{d}-{mode}-{dtype}[-causal].csv     (N + one ms column per provider, the
                       reference's file naming, src/bench.py:47, plus *_tflops columns; modes fwd, bwd, fwd_bwd — `bwd`
                       times O.backward(dO, retain_graph=True) alone as src/bench.py:91-99 does)
    gpurun_out/sweep.json                          (everything, incl. configs C2 / C3 / C4 and the fp32 rows)

N runs over 2^7 ... 2^15 like src/bench.py:11-12 (B = 8, H = 16; B = 4 at 2^14, B = 2 at 2^15 to bound the run).
The reference Triton kernels are imported from baseline/_ref/src (git-ignored copy of the UNMODIFIED reference
sources, made by tools/fetch_reference.sh in the dev container; absent -> column skipped).  Its 114-config autotune
list is trimmed to a handful of configs before import (README.md:29-31 warns about the search time), and the first
backward call is discarded (src/test_torch.py:23-28).
"""
from __future__ import annotations

import argparse
import csv
import json
import math
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from flash_attention_dlrs_b200 import FlashAttention, _native  # noqa: E402

DEV = torch.device("cuda", 0)


def flops(B, H, N, D, causal, mode):
    f = 4.0 * B * H * N * N * D * (0.5 if causal else 1.0)
    return {"fwd": f, "bwd": 2.5 * f, "fwd_bwd": 3.5 * f}[mode]


def timeit(fn, warmup=3, reps=20, budget_ms=250.0):
    """Mean ms per call, CUDA events on the launching stream.  `reps` is cut so that one measurement costs about
    `budget_ms` of GPU time (slow providers at long N would otherwise eat the GPU budget), never below 3."""
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    fn()
    b.record()
    torch.cuda.synchronize()
    one = max(a.elapsed_time(b), 1e-3)
    reps = int(max(3, min(reps, budget_ms / one)))
    for _ in range(min(warmup, reps)):
        fn()
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def load_reference_triton():
    src = os.path.join(ROOT, "baseline", "_ref", "src")
    if not os.path.isdir(src):
        return None, "baseline/_ref/src absent"
    sys.path.insert(0, src)
    # Unmodified, the reference does not compile under triton 3.6 (NameError: "Cannot access global variable ORDER
    # from within @jit'ed function", flash_attention_kernels.py:6-9 are plain-annotated globals); Triton's own escape
    # hatch keeps the sources untouched.
    os.environ.setdefault("TRITON_ALLOW_NON_CONSTEXPR_GLOBALS", "1")
    try:
        import triton
        import autotune_configs

        def short_list():
            C = triton.Config
            return [C({'B_r': 64, 'B_c': 64}, num_stages=2, num_warps=4), C({'B_r': 128, 'B_c': 64}, num_stages=2, num_warps=8),
                    C({'B_r': 64, 'B_c': 32}, num_stages=2, num_warps=4), C({'B_r': 32, 'B_c': 32}, num_stages=2, num_warps=4),
                    C({'B_r': 16, 'B_c': 16}, num_stages=2, num_warps=4),
                    C({'B_r': 64, 'B_c': 0}, num_stages=2, num_warps=4), C({'B_r': 128, 'B_c': 0}, num_stages=2, num_warps=4)]

        autotune_configs.get_autotune_config_cuda = short_list
        autotune_configs.SRAM = 200 * 1024  # the GA102 figure (autotune_configs.py:10) prunes everything useful on B200
        from flash_attention_torch import FlashAttention as RefFA  # noqa

        return RefFA, None
    except Exception as e:  # noqa
        return None, f"reference Triton import failed: {type(e).__name__}: {str(e)[:300]}"


def load_tutorial():
    """The vendored OpenAI tutorial kernel of the reference (flash_attention_openai_tutorial.py), unmodified."""
    src = os.path.join(ROOT, "baseline", "_ref", "src")
    if not os.path.isdir(src):
        return None, "baseline/_ref/src absent"
    if src not in sys.path:
        sys.path.insert(0, src)
    try:
        from flash_attention_openai_tutorial import _attention
        return _attention, None
    except Exception as e:  # noqa
        return None, f"tutorial import failed: {type(e).__name__}: {str(e)[:300]}"


def cpu_torch_ms(B, H, N, D, causal, scale, mode):
    """The reference's CPU path (fp32 SDPA + autograd, test_correctness.py:33,48) on 2 heads, scaled to B*H heads."""
    from oracle import attention_oracle as orc
    g = torch.Generator().manual_seed(42)
    Q, K, V, dO = (torch.randn(1, 2, N, D, generator=g) for _ in range(4))
    if mode == "fwd":
        orc.reference_sdpa(Q, K, V, scale, causal)
        t0 = time.perf_counter()
        orc.reference_sdpa(Q, K, V, scale, causal)
    else:
        orc.reference_sdpa_grads(Q, K, V, dO, scale, causal)
        t0 = time.perf_counter()
        orc.reference_sdpa_grads(Q, K, V, dO, scale, causal)
    return (time.perf_counter() - t0) * 1e3 * (B * H / 2.0)


def providers_for(dtype, causal, scale, want_ref, tutorial=None, N=0, deterministic_cudnn=False):
    out = {}

    def ours(q, k, v):
        return FlashAttention.apply(q, k, v, causal, scale)

    out["b200-cuda"] = ours
    try:
        from torch.nn.attention import SDPBackend, sdpa_kernel

        def sdpa(q, k, v):
            with sdpa_kernel(SDPBackend.FLASH_ATTENTION):
                return torch.nn.functional.scaled_dot_product_attention(q, k, v, scale=scale, is_causal=causal)

        out["torch-fa"] = sdpa

        def make(backend):
            def f(q, k, v):
                with sdpa_kernel(backend):
                    return torch.nn.functional.scaled_dot_product_attention(q, k, v, scale=scale, is_causal=causal)
            return f

        if dtype != torch.float32:
            out["torch-cudnn"] = make(SDPBackend.CUDNN_ATTENTION)
            if deterministic_cudnn:
                # the same backend asked for a deterministic backward (this library's backward always is)
                def cudnn_det(q, k, v):
                    with sdpa_kernel(SDPBackend.CUDNN_ATTENTION):
                        return torch.nn.functional.scaled_dot_product_attention(q, k, v, scale=scale, is_causal=causal)
                cudnn_det.deterministic = True
                out["torch-cudnn-det"] = cudnn_det
        out["torch-xformers"] = make(SDPBackend.EFFICIENT_ATTENTION)
        if N <= 4096:
            out["torch-math"] = make(SDPBackend.MATH)
        if dtype == torch.float32:
            del out["torch-fa"]          # the flash backend has no fp32 kernels
    except Exception:
        pass
    if tutorial is not None and dtype == torch.float16 and N >= 128:
        out["openai-tutorial"] = lambda q, k, v: tutorial.apply(q, k, v, causal, scale)
    try:
        if dtype == torch.float32:
            raise ImportError("flash_attn has no fp32 kernels")
        from flash_attn import flash_attn_func

        def dao(q, k, v):  # flash_attn wants (B, N, H, D)
            return flash_attn_func(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), softmax_scale=scale,
                                   causal=causal).transpose(1, 2)

        out["daolab-fa2"] = dao
    except Exception:
        pass
    if want_ref is not None and dtype in (torch.float16, torch.float32) and not causal and scale == 1.0:
        out["reference-triton"] = lambda q, k, v: want_ref.apply(q, k, v)
    return out


def bench_point(B, H, N, D, dtype, causal, scale, ref, modes=("fwd", "bwd", "fwd_bwd"), reps=20, tutorial=None,
                cpu=False, deterministic_cudnn=False):
    g = torch.Generator(device="cpu").manual_seed(42)
    Q, K, V, dO = (torch.randn(B, H, N, D, generator=g).to(dtype).to(DEV) for _ in range(4))
    row = {}
    for name, fn in providers_for(dtype, causal, scale, ref, tutorial, N, deterministic_cudnn).items():
        det = getattr(fn, "deterministic", False)
        if det:
            torch.use_deterministic_algorithms(True, warn_only=True)
        try:
            q, k, v = (t.clone().requires_grad_(True) for t in (Q, K, V))
            if "fwd" in modes:
                with torch.no_grad():
                    row[f"{name}_fwd_ms"] = timeit(lambda: fn(q, k, v), reps=reps)
            slow_ref = name == "reference-triton" and N > 4096   # keep the lock-based backward short
            if ("fwd_bwd" in modes or "bwd" in modes) and not slow_ref:
                if name == "reference-triton":  # first backward call of the reference is garbage: discard it
                    fn(q, k, v).backward(dO)

                def step():
                    q.grad = k.grad = v.grad = None
                    fn(q, k, v).backward(dO)

                if "fwd_bwd" in modes:
                    row[f"{name}_fwd_bwd_ms"] = timeit(step, reps=reps)
                if "bwd" in modes:   # the reference's own `bwd` mode (src/bench.py:91-99)
                    O = fn(q, k, v)

                    def bwd_only():
                        q.grad = k.grad = v.grad = None
                        O.backward(dO, retain_graph=True)

                    row[f"{name}_bwd_ms"] = timeit(bwd_only, reps=reps)
                    del O
        except Exception as e:  # noqa
            row[f"{name}_error"] = f"{type(e).__name__}: {str(e)[-400:]}"
        finally:
            if det:
                torch.use_deterministic_algorithms(False)
        torch.cuda.empty_cache()
    if cpu:
        for mode in modes:
            if mode != "bwd":
                row[f"cpu-torch_{mode}_ms"] = cpu_torch_ms(B, H, N, D, causal, scale, mode)
    for key in list(row):
        if key.endswith("_ms"):
            mode = "fwd_bwd" if key.endswith("fwd_bwd_ms") else ("bwd" if key.endswith("_bwd_ms") else "fwd")
            row[key.replace("_ms", "_tflops")] = flops(B, H, N, D, causal, mode) / (row[key] * 1e-3) / 1e12
    return row


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--no-ref", action="store_true")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu-torch column")
    ap.add_argument("--only", default="", help="comma list of sections: named,c5,fp32 (default: all)")
    args = ap.parse_args()
    only = set(x for x in args.only.split(",") if x) or {"named", "c5", "fp32"}
    os.makedirs(os.path.join(ROOT, "bench_out"), exist_ok=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    ref, ref_err = (None, "skipped") if args.no_ref else load_reference_triton()
    tut, tut_err = (None, "skipped") if args.no_ref else load_tutorial()
    results = {"reference_triton": "available" if ref is not None else ref_err,
               "openai_tutorial": "available" if tut is not None else tut_err,
               "cpu_threads": torch.get_num_threads(), "points": []}
    print("reference Triton:", results["reference_triton"], "| tutorial:", results["openai_tutorial"], flush=True)

    named = [
        ("C2", dict(B=4, H=16, N=4096, D=64, dtype=torch.float16, causal=False, scale=1.0)),          # reference semantics
        ("C2-scaled", dict(B=4, H=16, N=4096, D=64, dtype=torch.float16, causal=False, scale=0.125)),
        ("C3", dict(B=2, H=32, N=8192, D=128, dtype=torch.bfloat16, causal=True, scale=128 ** -0.5)),
        ("C4", dict(B=1, H=64, N=32768, D=128, dtype=torch.bfloat16, causal=True, scale=128 ** -0.5)),
    ]
    for tag, cfg in named:
        if (args.quick and tag == "C4") or "named" not in only:
            continue
        t0 = time.time()
        row = bench_point(ref=ref, reps=10 if tag == "C4" else 20, tutorial=tut, cpu=not args.no_cpu and tag != "C4",
                          deterministic_cudnn=True, **cfg)
        row.update(tag=tag, **{k: (str(v) if isinstance(v, torch.dtype) else v) for k, v in cfg.items()})
        results["points"].append(row)
        print(tag, {k: round(v, 3) for k, v in row.items() if isinstance(v, float)}, f"({time.time() - t0:.0f}s)", flush=True)

    # C5: the reference's sweep shape (src/bench.py:8-12: B=8, H=16, N = 2^7 .. 2^15) over N, D, causal; fp16 so that the
    # reference's kernels and the tutorial can run
    Ns = [512, 2048, 8192] if args.quick else [2 ** i for i in range(7, 16)]
    for D in (64, 128):
        for causal in (False, True):
            if "c5" not in only:
                continue
            for dtype, dname in ((torch.float16, "float16"),):
                rows = []
                for N in Ns:
                    Bq = 8 if N <= 8192 else (4 if N <= 16384 else 2)
                    scale = 1.0 if not causal else D ** -0.5   # non-causal at the reference's scale=1 so its kernel can run
                    row = bench_point(Bq, 16, N, D, dtype, causal, scale, ref, reps=20, tutorial=tut,
                                      cpu=not args.no_cpu and N in (512, 2048, 8192))
                    row.update(tag="C5", B=Bq, H=16, N=N, D=D, dtype=dname, causal=causal, scale=scale)
                    results["points"].append(row)
                    rows.append(row)
                    print("C5", D, causal, N, {k: round(v, 1) for k, v in row.items() if k.endswith("tflops")},
                          {k: v[:80] for k, v in row.items() if k.endswith("_error")}, flush=True)
                for mode in ("fwd", "bwd", "fwd_bwd"):
                    cols = sorted({k for r in rows for k in r if k.endswith(f"_{mode}_ms") or k.endswith(f"_{mode}_tflops")}
                                  - ({k for r in rows for k in r if k.endswith("_fwd_bwd_ms") or k.endswith("_fwd_bwd_tflops")}
                                     if mode == "bwd" else set()))
                    path = os.path.join(ROOT, "bench_out",
                                        f"fused-attention-B8-H16-d{D}-{mode}-{dname}{'-causal' if causal else ''}.csv")
                    with open(path, "w", newline="") as f:
                        w = csv.writer(f)
                        w.writerow(["N"] + cols)
                        for r in rows:
                            w.writerow([r["N"]] + [r.get(c, float("nan")) for c in cols])
    # fp32: the reference's primary tested dtype (src/test_correctness.py:13; DOT_PRECISION = "ieee", kernels.py:6) —
    # our SIMT fp32 path next to the reference's own fp32 Triton kernel, scale 1 non-causal (the reference's semantics)
    if "fp32" in only:
        for (Bq, Hq, N, D) in ((32, 32, 256, 128), (8, 16, 1024, 128), (8, 16, 4096, 128), (8, 16, 1024, 64), (8, 16, 4096, 64)):
            row = bench_point(Bq, Hq, N, D, torch.float32, False, 1.0, ref, reps=10, tutorial=None, cpu=False)
            row.update(tag="fp32", B=Bq, H=Hq, N=N, D=D, dtype="float32", causal=False, scale=1.0)
            results["points"].append(row)
            print("fp32", Bq, Hq, N, D, {k: round(v, 2) for k, v in row.items() if k.endswith("tflops")},
                  {k: v[:80] for k, v in row.items() if k.endswith("_error")}, flush=True)
    with open(os.path.join(ROOT, "gpurun_out", "sweep.json"), "w") as f:
        json.dump(results, f, indent=1)
    print("wrote gpurun_out/sweep.json")


if __name__ == "__main__":
    main()
