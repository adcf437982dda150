#!/usr/bin/env python
"""Sequence-length sweep and the other BASELINE.json configs, next to the baselines the reference's own bench lines
up (src/bench.py:34-55): the reference Triton kernel (fp16, scale 1, non-causal only — where it compiles), torch
SDPA flash, DaoLab flash_attn.  Not part of the driver contract (that is bench.py); writes

    bench_out/fused-attention-B{B}-H{H}-d{d}-{mode}-{dtype}[-causal].csv     (N + one ms column per provider, the
                                                  reference's file naming, src/bench.py:47, plus *_tflops columns)
    gpurun_out/sweep.json                          (everything, incl. configs C2 / C3 / C4)

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


def timeit(fn, warmup=5, reps=20):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
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


def providers_for(dtype, causal, scale, want_ref):
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
    except Exception:
        pass
    try:
        from flash_attn import flash_attn_func

        def dao(q, k, v):  # flash_attn wants (B, N, H, D)
            return flash_attn_func(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), softmax_scale=scale,
                                   causal=causal).transpose(1, 2)

        out["daolab-fa2"] = dao
    except Exception:
        pass
    if want_ref is not None and dtype == torch.float16 and not causal and scale == 1.0:
        out["reference-triton"] = lambda q, k, v: want_ref.apply(q, k, v)
    return out


def bench_point(B, H, N, D, dtype, causal, scale, ref, modes=("fwd", "fwd_bwd"), reps=20):
    g = torch.Generator(device="cpu").manual_seed(42)
    Q, K, V, dO = (torch.randn(B, H, N, D, generator=g).to(dtype).to(DEV) for _ in range(4))
    row = {}
    for name, fn in providers_for(dtype, causal, scale, ref).items():
        try:
            q, k, v = (t.clone().requires_grad_(True) for t in (Q, K, V))
            if "fwd" in modes:
                with torch.no_grad():
                    row[f"{name}_fwd_ms"] = timeit(lambda: fn(q, k, v), reps=reps)
            if "fwd_bwd" in modes and not (name == "reference-triton" and N > 4096):  # keep the lock-based bwd short
                if name == "reference-triton":  # first backward call of the reference is garbage: discard it
                    fn(q, k, v).backward(dO)

                def step():
                    q.grad = k.grad = v.grad = None
                    fn(q, k, v).backward(dO)

                row[f"{name}_fwd_bwd_ms"] = timeit(step, reps=reps)
        except Exception as e:  # noqa
            row[f"{name}_error"] = f"{type(e).__name__}: {str(e)[-400:]}"
    for key in list(row):
        if key.endswith("_ms"):
            mode = "fwd_bwd" if key.endswith("fwd_bwd_ms") else "fwd"
            row[key.replace("_ms", "_tflops")] = flops(B, H, N, D, causal, mode) / (row[key] * 1e-3) / 1e12
    return row


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--no-ref", action="store_true")
    args = ap.parse_args()
    os.makedirs(os.path.join(ROOT, "bench_out"), exist_ok=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    ref, ref_err = (None, "skipped") if args.no_ref else load_reference_triton()
    results = {"reference_triton": "available" if ref is not None else ref_err, "points": []}
    print("reference Triton:", results["reference_triton"], flush=True)

    named = [
        ("C2", dict(B=4, H=16, N=4096, D=64, dtype=torch.float16, causal=False, scale=1.0)),          # reference semantics
        ("C2-scaled", dict(B=4, H=16, N=4096, D=64, dtype=torch.float16, causal=False, scale=0.125)),
        ("C3", dict(B=2, H=32, N=8192, D=128, dtype=torch.bfloat16, causal=True, scale=128 ** -0.5)),
        ("C4", dict(B=1, H=64, N=32768, D=128, dtype=torch.bfloat16, causal=True, scale=128 ** -0.5)),
    ]
    for tag, cfg in named:
        if args.quick and tag == "C4":
            continue
        t0 = time.time()
        row = bench_point(ref=ref, reps=10 if tag == "C4" else 20, **cfg)
        row.update(tag=tag, **{k: (str(v) if isinstance(v, torch.dtype) else v) for k, v in cfg.items()})
        results["points"].append(row)
        print(tag, {k: round(v, 3) for k, v in row.items() if isinstance(v, float)}, f"({time.time() - t0:.0f}s)", flush=True)

    # C5: the reference's sweep shape (src/bench.py:8-12: B=8, H=16) over N, D, causal; fp16 so the reference can run
    Ns = [512, 2048, 8192] if args.quick else [512, 1024, 2048, 4096, 8192, 16384]
    for D in (64, 128):
        for causal in (False, True):
            for dtype, dname in ((torch.float16, "float16"),):
                rows = []
                for N in Ns:
                    Bq = 8 if N <= 8192 else 4
                    scale = 1.0 if not causal else D ** -0.5   # non-causal at the reference's scale=1 so its kernel can run
                    row = bench_point(Bq, 16, N, D, dtype, causal, scale, ref, reps=10)
                    row.update(tag="C5", B=Bq, H=16, N=N, D=D, dtype=dname, causal=causal, scale=scale)
                    results["points"].append(row)
                    rows.append(row)
                    print("C5", D, causal, N, {k: round(v, 1) for k, v in row.items() if k.endswith("tflops")}, flush=True)
                for mode in ("fwd", "fwd_bwd"):
                    cols = sorted({k for r in rows for k in r if k.endswith(f"_{mode}_ms") or k.endswith(f"_{mode}_tflops")})
                    path = os.path.join(ROOT, "bench_out",
                                        f"fused-attention-B8-H16-d{D}-{mode}-{dname}{'-causal' if causal else ''}.csv")
                    with open(path, "w", newline="") as f:
                        w = csv.writer(f)
                        w.writerow(["N"] + cols)
                        for r in rows:
                            w.writerow([r["N"]] + [r.get(c, float("nan")) for c in cols])
    with open(os.path.join(ROOT, "gpurun_out", "sweep.json"), "w") as f:
        json.dump(results, f, indent=1)
    print("wrote gpurun_out/sweep.json")


if __name__ == "__main__":
    main()
