"""Forward kernel A/B: FA_FWD_PAIR=0 (single CTA) vs 1 (CTA pairs).  Prints, per shape, the time and a checksum of the raw
bits of O and L; the two runs must print the same checksums (the pair kernel changes who feeds the tensor cores, not the
arithmetic) and each run is compared against torch SDPA (math, fp32) as a sanity bound."""
import hashlib, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flash_attention_dlrs_b200 import _native
dev = torch.device("cuda", 0)
def t(fn, reps=20):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
mode = "pair=%s,w16=%s,duo=%s" % (os.environ.get("FA_FWD_PAIR", "d"), os.environ.get("FA_FWD_W16", "d"), os.environ.get("FA_FWD_DUO", "d"))
save = os.environ.get("FA_PROBE_SAVE")      # directory: store O / L of every shape
cmp_ = os.environ.get("FA_PROBE_COMPARE")   # directory: compare with what another mode stored there
small = [(2, 3, n, d, c, dt) for d in (128, 64) for n in (1, 100, 128, 200, 384, 512, 513, 640, 1024, 1100, 2048) for c in (False, True)
         for dt in (torch.bfloat16, torch.float16)]
big = [(2, 32, 8192, 128, True, torch.bfloat16), (2, 32, 8192, 128, False, torch.bfloat16), (1, 16, 32768, 128, True, torch.bfloat16),
       (8, 16, 512, 128, False, torch.float16), (8, 16, 1024, 128, True, torch.float16),
       (4, 16, 4096, 64, False, torch.float16), (8, 16, 512, 64, True, torch.float16), (8, 16, 8192, 64, True, torch.float16)]
for (B, H, N, D, causal, dt) in small + big:
    g = torch.Generator().manual_seed(N + D)
    Q, K, V = (torch.randn(B, H, N, D, generator=g).to(dt).to(dev) for _ in range(3))
    sc = D ** -0.5
    O, L = _native.forward(Q, K, V, causal, sc)
    torch.cuda.synchronize()
    hs = hashlib.sha1(O.cpu().view(torch.int16).numpy().tobytes() + L.cpu().numpy().tobytes()).hexdigest()[:10]
    err = ""
    if N <= 2048:
        with torch.nn.attention.sdpa_kernel(torch.nn.attention.SDPBackend.MATH):
            ref = torch.nn.functional.scaled_dot_product_attention(Q.float(), K.float(), V.float(), scale=sc, is_causal=causal)
        err = " max|O-ref| %.2e" % (O.float() - ref).abs().max().item()
    tag = "B%dH%dN%dD%dc%d%s" % (B, H, N, D, causal, str(dt)[6:])
    if save:
        os.makedirs(save, exist_ok=True)
        torch.save((O.cpu(), L.cpu()), os.path.join(save, tag + ".pt"))
    if cmp_ and os.path.exists(os.path.join(cmp_, tag + ".pt")):
        O2, L2 = torch.load(os.path.join(cmp_, tag + ".pt"))
        err += " | vs other mode: max|dO| %.2e max|dL| %.2e%s" % ((O.cpu().float() - O2.float()).abs().max().item(),
                                                                   (L.cpu() - L2).abs().max().item(), " (bits equal)" if torch.equal(O.cpu(), O2) and torch.equal(L.cpu(), L2) else "")
    ms = t(lambda: _native.forward(Q, K, V, causal, sc)) if N >= 512 else 0.0
    fl = 4.0 * B * H * N * N * D * (0.5 if causal else 1.0)
    print("%s B%d H%d N%d D%d c%d %s  %s  %.3f ms %.0f TFLOP/s%s" % (mode, B, H, N, D, causal, str(dt)[6:], hs, ms, fl / max(ms, 1e-9) / 1e9 if ms else 0, err), flush=True)
