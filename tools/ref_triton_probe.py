"""Does the UNMODIFIED reference Triton forward compile and run on this box (triton version printed)?  Records the
full error text otherwise.  Needs baseline/_ref/src (tools/fetch_reference.sh)."""
import os, sys, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
if len(sys.argv) > 1 and sys.argv[1] == "allow-globals":
    os.environ["TRITON_ALLOW_NON_CONSTEXPR_GLOBALS"] = "1"
else:
    os.environ["TRITON_ALLOW_NON_CONSTEXPR_GLOBALS"] = "0"
import torch, triton
print("triton", triton.__version__, "torch", torch.__version__, torch.cuda.get_device_name(0))
from bench_sweep import load_reference_triton
ref, err = load_reference_triton()
print("import:", "ok" if ref is not None else err)
if ref is not None:
    Q, K, V = (torch.randn(1, 2, 256, 64, device="cuda", dtype=torch.float16) for _ in range(3))
    try:
        O = ref.apply(Q, K, V)
        torch.cuda.synchronize()
        O_ref = torch.nn.functional.scaled_dot_product_attention(Q, K, V, scale=1)
        print("reference Triton fwd ran; max|O - sdpa| =", (O.float() - O_ref.float()).abs().max().item())
    except Exception as e:
        msg = "".join(traceback.format_exception_only(type(e), e))
        print("reference Triton fwd FAILED:\n" + msg[-3000:])
