#!/usr/bin/env python
"""Markdown tables of a tools/bench_sweep.py result (gpurun_out/sweep.json or profiles/rNN_sweep.json): TFLOP/s per provider.

    python tools/render_sweep.py profiles/r02_sweep.json > /tmp/tables.md
"""
import json
import sys

PROVIDERS = ["b200-cuda", "torch-cudnn", "reference-triton", "openai-tutorial", "daolab-fa2", "torch-fa", "torch-xformers",
             "torch-math", "cpu-torch"]
HEAD = {"b200-cuda": "ours", "torch-cudnn": "torch SDPA cuDNN", "reference-triton": "reference Triton", "openai-tutorial": "tutorial Triton",
        "daolab-fa2": "flash_attn 2.8", "torch-fa": "SDPA flash", "torch-xformers": "SDPA efficient", "torch-math": "SDPA math",
        "cpu-torch": "CPU torch (scaled)"}


def fmt(v):
    if not isinstance(v, (int, float)) or v != v:
        return "—"
    return f"{v:.0f}" if v >= 100 else (f"{v:.1f}" if v >= 1 else f"{v:.2f}")


def table(rows, label, modes, provs):
    out = ["| " + " | ".join([label, "mode"] + [HEAD[p] for p in provs]) + " |", "|" + "---|" * (2 + len(provs))]
    for name, r in rows:
        for m in modes:
            out.append("| " + " | ".join([name, m] + [fmt(r.get(f"{p}_{m}_tflops")) for p in provs]) + " |")
    return "\n".join(out)


def main():
    d = json.load(open(sys.argv[1]))
    pts = d["points"]
    print(f"reference Triton: {d['reference_triton']}; tutorial: {d['openai_tutorial']}; CPU threads: {d['cpu_threads']}\n")
    named = [(f"{r['tag']} {r['dtype'].replace('torch.', '')} B{r['B']} H{r['H']} N{r['N']} D{r['D']} {'causal' if r['causal'] else 'non-causal'} scale {r['scale']:.3g}", r)
             for r in pts if r["tag"] in ("C2", "C2-scaled", "C3", "C4")]
    print("### BASELINE configs (TFLOP/s algorithmic)\n")
    print(table(named, "config", ("fwd", "bwd", "fwd_bwd"), PROVIDERS))
    for D in (64, 128):
        for causal in (False, True):
            rows = [(f"N={r['N']} (B{r['B']})", r) for r in pts if r["tag"] == "C5" and r["D"] == D and r["causal"] == causal]
            if rows:
                print(f"\n### C5 sweep fp16 H16 D={D} {'causal' if causal else 'non-causal (scale 1, the reference semantics)'} — fwd_bwd and fwd TFLOP/s\n")
                print(table(rows, "N", ("fwd", "fwd_bwd"), [p for p in PROVIDERS if p != "torch-math"]))
    rows = [(f"B{r['B']} H{r['H']} N{r['N']} D{r['D']}", r) for r in pts if r["tag"] == "fp32"]
    if rows:
        print("\n### float32 (the reference's tested dtype), scale 1 non-causal — TFLOP/s\n")
        print(table(rows, "shape", ("fwd", "fwd_bwd"), ["b200-cuda", "reference-triton", "torch-xformers", "torch-math"]))


if __name__ == "__main__":
    main()
