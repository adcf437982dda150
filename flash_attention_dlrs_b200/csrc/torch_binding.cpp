// torch_binding.cpp — C++ autograd node over the C ABI (include/fa_b200.h) for the plain attention call.
//
// FlashAttention.apply(Q, K, V[, causal, softmax_scale]) in the reference is a Python torch.autograd.Function around
// Triton launches (flash_attention_torch.py:21-158).  The Python Function of this package does the same around
// libfa_b200.so and stays the general path (seqlens, dropout, masks, padded head sizes, FP8); for the plain call — the only
// one the reference has — this node runs the same two C-ABI calls without the interpreter: no ctypes marshalling and no GIL
// hand-over when the autograd engine's device thread runs the backward.  Below N ~ 1024 a fwd+bwd step is bound by that
// host path (tools/small_sweep_probe.py), not by the kernels.
//
// Built in-tree as flash_attention_dlrs_b200/_fa_torch*.so (see _lib.build_torch_binding), linked against libfa_b200.so
// next to it ($ORIGIN).  Same results bit for bit as the Python path: it calls the same entry points with the same arguments.
#include <torch/extension.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>

#include "fa_b200.h"

namespace {

int dtype_code(at::ScalarType t) {
  switch (t) {
    case at::kHalf: return FA_DTYPE_F16;
    case at::kBFloat16: return FA_DTYPE_BF16;
    case at::kFloat: return FA_DTYPE_F32;
    default: return -1;
  }
}

// A view the kernels can address (the checks of _native._kernel_ready): unit inner stride, 16-byte aligned base and
// outer strides, no broadcast / reversed dimensions; anything else is copied.
at::Tensor kernel_ready(const at::Tensor& t) {
  const int64_t gran = 16 / (int64_t)t.element_size();
  const auto sz = t.sizes();
  const auto st = t.strides();
  bool ok = st[3] == 1 && reinterpret_cast<uintptr_t>(t.data_ptr()) % 16 == 0 && st[2] % gran == 0 && st[2] > 0;
  ok = ok && (sz[1] == 1 || (st[1] % gran == 0 && st[1] > 0)) && (sz[0] == 1 || (st[0] % gran == 0 && st[0] > 0));
  ok = ok && (sz[2] == 1 || st[2] >= sz[3]);
  return ok ? t : t.contiguous();
}

void strides4(const at::Tensor& t, int64_t (&s)[4]) {
  for (int i = 0; i < 4; ++i) s[i] = t.stride(i);
}

void check(int rc, const char* what) {
  TORCH_CHECK(rc == 0, what, " failed (code ", rc, "): ", fa_last_error());
}

struct FlashAttentionNode : public torch::autograd::Function<FlashAttentionNode> {
  static at::Tensor forward(torch::autograd::AutogradContext* ctx, const at::Tensor& Q, const at::Tensor& K,
                            const at::Tensor& V, bool causal, double softmax_scale) {
    const c10::cuda::CUDAGuard guard(Q.device());
    const auto q = kernel_ready(Q), k = kernel_ready(K), v = kernel_ready(V);
    const int B = (int)Q.size(0), H = (int)Q.size(1), N = (int)Q.size(2), D = (int)Q.size(3);
    auto O = at::empty({B, H, N, D}, Q.options());
    auto L = at::empty({B, H, N}, Q.options().dtype(at::kFloat));
    int64_t qs[4], ks[4], vs[4], os[4];
    strides4(q, qs), strides4(k, ks), strides4(v, vs), strides4(O, os);
    check(fa_fwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), O.data_ptr(), L.data_ptr<float>(), B, H, N, D, qs, ks, vs, os,
                 dtype_code(Q.scalar_type()), (float)softmax_scale, causal ? 1 : 0,
                 c10::cuda::getCurrentCUDAStream().stream()),
          "fa_fwd");
    // same saved set as the reference (flash_attention_torch.py:77)
    ctx->save_for_backward({Q, K, V, O, L});
    ctx->saved_data["causal"] = causal;
    ctx->saved_data["scale"] = softmax_scale;
    return O;
  }

  static torch::autograd::variable_list backward(torch::autograd::AutogradContext* ctx,
                                                 torch::autograd::variable_list grads) {
    const auto saved = ctx->get_saved_variables();
    const at::Tensor &Q = saved[0], &K = saved[1], &V = saved[2], &O = saved[3], &L = saved[4];
    TORCH_CHECK_VALUE(grads[0].scalar_type() == Q.scalar_type(), "dO must have same dtype as inputs");
    const c10::cuda::CUDAGuard guard(Q.device());
    const auto q = kernel_ready(Q), k = kernel_ready(K), v = kernel_ready(V), o = kernel_ready(O);
    const auto dout = kernel_ready(grads[0]);
    const int B = (int)Q.size(0), H = (int)Q.size(1), N = (int)Q.size(2), D = (int)Q.size(3);
    const int code = dtype_code(Q.scalar_type());
    void* stream = c10::cuda::getCurrentCUDAStream().stream();
    auto delta = at::empty({B, H, N}, Q.options().dtype(at::kFloat));
    int64_t qs[4], ks[4], vs[4], os[4], ds[4], gs[4];
    strides4(q, qs), strides4(k, ks), strides4(v, vs), strides4(o, os), strides4(dout, ds);
    check(fa_bwd_preprocess(o.data_ptr(), dout.data_ptr(), delta.data_ptr<float>(), B, H, N, D, os, ds, code, stream),
          "fa_bwd_preprocess");
    // three independent allocations: a caller that keeps only one gradient alive must not pin the other two
    auto dQ = at::empty({B, H, N, D}, Q.options()), dK = at::empty({B, H, N, D}, Q.options()),
         dV = at::empty({B, H, N, D}, Q.options());
    strides4(dQ, gs);
    check(fa_bwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), dout.data_ptr(), L.data_ptr<float>(), delta.data_ptr<float>(),
                 dQ.data_ptr(), dK.data_ptr(), dV.data_ptr(), nullptr, 0, B, H, N, D, qs, ks, vs, ds, gs, gs, gs, code,
                 (float)ctx->saved_data["scale"].toDouble(), ctx->saved_data["causal"].toBool() ? 1 : 0, stream),
          "fa_bwd");
    return {dQ, dK, dV, at::Tensor(), at::Tensor()};
  }
};

// True when the plain call can take this path: 4-D same-shape same-dtype CUDA tensors on one device, a dtype and head
// size the kernels run at without padding, something to compute.  Everything else goes through the Python Function,
// which also owns the error messages.
bool supported(const at::Tensor& Q, const at::Tensor& K, const at::Tensor& V) {
  if (!Q.is_cuda() || Q.dim() != 4 || K.sizes() != Q.sizes() || V.sizes() != Q.sizes()) return false;
  if (K.device() != Q.device() || V.device() != Q.device()) return false;
  if (K.scalar_type() != Q.scalar_type() || V.scalar_type() != Q.scalar_type()) return false;
  const int code = dtype_code(Q.scalar_type());
  const int64_t d = Q.size(3);
  if (code < 0 || Q.numel() == 0) return false;
  if (code == FA_DTYPE_F32) return d == 16 || d == 32 || d == 64 || d == 128;
  return d == 64 || d == 128;
}

at::Tensor flash_attention(const at::Tensor& Q, const at::Tensor& K, const at::Tensor& V, bool causal,
                           double softmax_scale) {
  return FlashAttentionNode::apply(Q, K, V, causal, softmax_scale);
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("supported", &supported, "can the plain call (no seqlens / dropout / mask) take the C++ autograd node?");
  m.def("flash_attention", &flash_attention, "O = attention(Q, K, V) through the C++ autograd node over libfa_b200.so");
  m.def("abi_version", []() { return fa_version(); });
}
