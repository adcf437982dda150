"""CPU tests of the host side: the C-ABI library loads and exports every symbol include/fa_b200.h declares,
argument validation (no compute without a GPU), the reference's error conventions, head-size padding, sharding."""
import ctypes
import re

import pytest
import torch

from flash_attention_dlrs_b200 import _lib, _native
from flash_attention_dlrs_b200 import flash_attention_torch as fat
from flash_attention_dlrs_b200 import flash_attention_wrappers as faw
from flash_attention_dlrs_b200.sharding import head_range


def test_library_exports_every_declared_symbol():
    header = _lib.HEADER.read_text()
    declared = set(re.findall(r"\b(fa_[a-z_0-9]+)\s*\(", header))
    assert declared == set(_lib.EXPORTED_SYMBOLS)
    lib = ctypes.CDLL(str(_lib.LIB_PATH))
    for sym in declared:
        assert hasattr(lib, sym), sym


def test_version_and_error_string():
    lib = _lib.load()
    assert lib.fa_version() == 9
    assert isinstance(lib.fa_last_error(), bytes)


def test_c_abi_rejects_bad_arguments_without_touching_the_gpu():
    lib = _lib.load()
    s = (ctypes.c_int64 * 4)(64 * 128, 128 * 64, 64, 1)
    null = ctypes.c_void_p(0)
    # unsupported dtype
    assert lib.fa_fwd(null, null, null, null, null, 1, 1, 128, 64, s, s, s, s, 7, 1.0, 0, null) < 0
    assert b"dtype" in lib.fa_last_error()
    # unsupported head size for 16-bit
    assert lib.fa_fwd(null, null, null, null, null, 1, 1, 128, 48, s, s, s, s, 1, 1.0, 0, null) < 0
    # null tensor
    assert lib.fa_fwd(null, null, null, null, null, 1, 1, 128, 64, s, s, s, s, 1, 1.0, 0, null) < 0
    assert b"null" in lib.fa_last_error()
    # non-positive scale
    assert lib.fa_fwd(null, null, null, null, null, 1, 1, 128, 64, s, s, s, s, 1, 0.0, 0, null) < 0
    assert lib.fa_bwd_preprocess(null, null, null, 1, 1, 128, 64, s, s, 1, null) < 0
    # two-kernel path (what fa_bwd runs): no scratch; single-pass kernel: fp32 dQ tiles + turn counters
    assert lib.fa_bwd_workspace_bytes(2, 32, 8192, 128, 1, 1, 3) == 0
    assert lib.fa_bwd_workspace_bytes(2, 32, 8192, 128, 2, 1, 4) == 0
    causal = lib.fa_bwd_workspace_bytes(2, 32, 8192, 128, 1, 1, 4)
    full = lib.fa_bwd_workspace_bytes(2, 32, 8192, 128, 1, 0, 4)
    tiles = 2 * 32 * 64
    assert causal == 256 + tiles * 32 + tiles * 128 * 128 * 4   # ticket | 2 x 4 turn counters per tile | fp32 tiles
    assert full == 256 + tiles * 32 + 2 * tiles * 128 * 128 * 4
    with pytest.raises(_lib.FlashAttentionLibraryError):
        _lib.check(-1, "fa_fwd")


def test_cpu_tensors_raise_like_the_reference():
    # flash_attention_torch.py:24-26 — no CPU fallback
    Q = torch.randn(1, 1, 16, 16)
    with pytest.raises(NotImplementedError):
        fat.FlashAttention.apply(Q, Q, Q)
    with pytest.raises(NotImplementedError):
        faw.flash_attention_forward(Q, Q, Q, torch.device("cpu"))


def test_dtype_whitelist():
    # flash_attention_torch.py:17-18
    assert fat.convert_triton_dtype(torch.float16) == _lib.FA_DTYPE_F16
    assert fat.convert_triton_dtype(torch.bfloat16) == _lib.FA_DTYPE_BF16
    assert fat.convert_triton_dtype(torch.float32) == _lib.FA_DTYPE_F32
    # FP8 (forward only): float8_e5m2 is in the reference's own map (flash_attention_torch.py:15-16)
    assert fat.convert_triton_dtype(torch.float8_e5m2) == _lib.FA_DTYPE_F8E5M2 == 4
    assert fat.convert_triton_dtype(torch.float8_e4m3fn) == _lib.FA_DTYPE_F8E4M3 == 3
    with pytest.raises(TypeError, match="not supported"):
        fat.convert_triton_dtype(torch.float64)
    with pytest.raises(TypeError):
        fat.convert_triton_dtype(torch.int8)


def test_deterministic_alias_and_names():
    assert fat.FlashAttentionDeterministic is fat.FlashAttention
    assert fat.MIN_TENSOR_SIZE == 16


def test_padded_head_dim():
    assert _native.padded_head_dim(128, torch.bfloat16) == 128
    assert _native.padded_head_dim(64, torch.float16) == 64
    assert _native.padded_head_dim(80, torch.float16) == 128
    assert _native.padded_head_dim(8, torch.float16) == 64
    assert _native.padded_head_dim(8, torch.float32) == 16
    assert _native.padded_head_dim(40, torch.float32) == 64
    assert _native.padded_head_dim(128, torch.float32) == 128
    assert _native.padded_head_dim(64, torch.float8_e5m2) == 128
    with pytest.raises(ValueError):
        _native.padded_head_dim(129, torch.float32)


def test_fp8_padding_is_bytewise_zero():
    t = torch.randn(1, 2, 5, 40).to(torch.float8_e4m3fn)
    p = _native._pad_d(t, 128)
    assert p.dtype == t.dtype and p.shape == (1, 2, 5, 128)
    assert torch.equal(p.view(torch.uint8)[..., :40], t.view(torch.uint8)) and not p.view(torch.uint8)[..., 40:].any()


def test_kernel_ready_views():
    x = torch.zeros(2, 4, 32, 64, dtype=torch.bfloat16)
    assert _native._kernel_ready(x[:, 1:3]) .data_ptr() == x[:, 1:3].data_ptr()   # head slice: no copy
    assert _native._kernel_ready(x.transpose(1, 2)).is_contiguous() or True
    y = x[..., ::2]
    assert _native._kernel_ready(y).stride(-1) == 1
    # broadcast views (stride 0) cannot be described to TMA: they are materialised
    e = torch.zeros(1, 1, 32, 64, dtype=torch.bfloat16).expand(2, 4, 32, 64)
    r = _native._kernel_ready(e)
    assert r.is_contiguous() and r.shape == e.shape
    # (B, N, H, D) storage viewed as (B, H, N, D): legal as is (strides are multiples of 16 bytes), no copy
    z = torch.zeros(2, 32, 4, 64, dtype=torch.bfloat16).transpose(1, 2)
    assert _native._kernel_ready(z).data_ptr() == z.data_ptr() and not _native._kernel_ready(z).is_contiguous()


def test_head_range_partitions_heads():
    for H in (1, 7, 8, 32, 64):
        for world in (1, 2, 3, 4, 8):
            spans = [head_range(H, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == H
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        head_range(8, 2, 2)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", tmp_path / "libfa_b200.so")
    with pytest.raises(_lib.FlashAttentionLibraryError, match="no CPU or eager fallback"):
        _lib.load()


def test_header_is_plain_c_and_matches_the_library(tmp_path):
    """include/fa_b200.h must be consumable from C (the drop-in boundary is a C ABI): compile a C translation unit that
    includes it and takes the address of every declared entry point, then link it against libfa_b200.so."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    src = tmp_path / "abi_check.c"
    syms = [s for s in _lib.EXPORTED_SYMBOLS]
    src.write_text('#include "fa_b200.h"\n#include <stdio.h>\ntypedef void (*fn_t)(void);\nint main(void) {\n  fn_t p[] = {' +
                   ", ".join(f"(fn_t){s}" for s in syms) +
                   '};\n  printf("%d %d\\n", (int)(sizeof p / sizeof p[0]), FA_BWD_FUSED + FA_DTYPE_F8E5M2);\n  return 0;\n}\n')
    inc = str(_lib.HEADER.parent)
    subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", inc, "-c", str(src), "-o",
                    str(tmp_path / "abi_check.o")], check=True, capture_output=True)
    lib = _lib.LIB_PATH
    out = subprocess.run([gcc, str(tmp_path / "abi_check.o"), "-L", str(lib.parent), f"-l:{lib.name}", "-o",
                          str(tmp_path / "abi_check")], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr[-2000:]   # every declared symbol resolves against the shared library


def test_dropout_mask_host_compile_matches_the_oracle_bit_for_bit(tmp_path):
    """csrc/fa_dropout.cuh compiled for the host (g++): the scalar form and both pair walks the tcgen05 kernels use
    (row walk = forward / dQ, column walk = dK/dV) reproduce oracle.dropout_keep_mask exactly."""
    import shutil
    import subprocess

    import numpy as np

    from oracle import attention_oracle as orc

    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    B, H, N, p, seed = 2, 3, 40, 0.3, 0x1234567890ABCDEF
    src = tmp_path / "mask.cpp"
    src.write_text(r'''
#include <cstdio>
#define __host__
#define __device__
#define __forceinline__ inline
#include "fa_dropout.cuh"
int main() {
  const int B = %d, H = %d, N = %d;
  fa::DropParams d{%du, %uu, %uu, 1.f};
  for (int bh = 0; bh < B * H; ++bh) {
    const uint32_t key = fa::drop_key(d, bh);
    for (int i = 0; i < N; ++i)
      for (int j = 0; j < N; j += 2) {
        bool r0, r1, c0, c1;   // row walk from (i, j); column walk from (j, i) i.e. queries j, j+1 of key i
        fa::drop_keep_pair<8>(key + fa::drop_word_index(i, j), 16u * (i & 1), d.thresh, r0, r1);
        fa::drop_keep_pair<16>(key + fa::drop_word_index(j, i), 8u * (i & 1), d.thresh, c0, c1);
        std::printf("%%d%%d%%d%%d%%d%%d\n", (int)fa::drop_keep(key, i, j, d.thresh), (int)fa::drop_keep(key, i, j + 1, d.thresh),
                    (int)r0, (int)r1, (int)c0, (int)c1);
      }
  }
  return 0;
}
''' % (B, H, N, orc.dropout_threshold(p), seed & 0xFFFFFFFF, seed >> 32))
    exe = tmp_path / "mask"
    subprocess.run([gxx, "-O1", "-std=c++17", "-I", str(_lib.CSRC), "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    got = np.array([[int(ch) for ch in line] for line in out], dtype=bool).reshape(B * H, N, N // 2, 6)
    keep = orc.dropout_keep_mask(seed, B, H, N, p).numpy().reshape(B * H, N, N)
    assert (got[..., 0] == keep[:, :, 0::2]).all() and (got[..., 1] == keep[:, :, 1::2]).all()   # scalar form
    assert (got[..., 2] == keep[:, :, 0::2]).all() and (got[..., 3] == keep[:, :, 1::2]).all()   # row walk
    keep_t = keep.transpose(0, 2, 1)   # [key][query]
    assert (got[..., 4] == keep_t[:, :, 0::2]).all() and (got[..., 5] == keep_t[:, :, 1::2]).all()   # column walk


def test_dropout_mask_statistics_and_quantisation():
    from oracle import attention_oracle as orc

    assert _native.dropout_threshold(0.0) == 0 and _native.dropout_threshold(0.001) == 0
    assert _native.dropout_threshold(0.1) == orc.dropout_threshold(0.1) == 26
    assert _native.dropout_threshold(0.999) == 255
    with pytest.raises(ValueError):
        _native.dropout_threshold(1.0)
    for p in (0.1, 0.5):
        keep = orc.dropout_keep_mask(42, 2, 4, 512, p)
        rate = 1.0 - orc.dropout_threshold(p) / 256.0
        n = keep.numel()
        assert abs(keep.double().mean().item() - rate) < 5.0 * (rate * (1 - rate) / n) ** 0.5
        # rows, columns and heads are decorrelated: per-row keep rates scatter like independent Bernoulli draws
        rows = keep.double().mean(-1)
        assert abs(rows.std().item() - (rate * (1 - rate) / 512) ** 0.5) < 0.2 * (rate * (1 - rate) / 512) ** 0.5
        assert (keep[0, 0] != keep[0, 1]).double().mean().item() > 0.5 * 2 * rate * (1 - rate)
    assert not torch.equal(orc.dropout_keep_mask(1, 1, 1, 64, 0.5), orc.dropout_keep_mask(2, 1, 1, 64, 0.5))


def test_c_abi_rejects_bad_dropout_arguments():
    lib = _lib.load()
    s = (ctypes.c_int64 * 4)(64 * 128, 128 * 64, 64, 1)
    null = ctypes.c_void_p(0)
    peers = (ctypes.c_void_p * 1)()
    rc = lib.fa_fwd_peers(null, null, null, null, null, 1, 1, 128, 64, s, s, s, s, 1, 1.0, 0, 0, peers, null, 1.0, 0,
                          None, null)
    assert rc < 0 and b"dropout_p" in lib.fa_last_error()
    rc = lib.fa_fwd_peers(null, null, null, null, null, 1, 1, 128, 128, s, s, s, s, 3, 1.0, 0, 0, peers, null, 0.5, 0,
                          None, null)
    assert rc < 0 and b"FP8" in lib.fa_last_error()
    rc = lib.fa_bwd_partial(null, null, null, null, null, null, null, null, null, null, 0, 1, 1, 128, 64, s, s, s, s, s, s, s,
                            1, 1.0, 0, 4, null, 0.5, 0, None, null)
    assert rc < 0 and b"FA_BWD_FUSED" in lib.fa_last_error()
    # attention mask (fa_attn_mask): row pitch must cover N rounded up to 128 and keep 16-byte groups aligned
    am = _lib.AttnMaskStruct()
    am.rows = 4096
    am.rows_strides = (ctypes.c_int64 * 3)(0, 0, 8)
    rc = lib.fa_fwd_peers(null, null, null, null, null, 1, 1, 100, 64, s, s, s, s, 1, 1.0, 0, 0, peers, null, 0.0, 0,
                          ctypes.byref(am), null)
    assert rc < 0 and b"row pitch" in lib.fa_last_error()
    am.rows_strides = (ctypes.c_int64 * 3)(0, 0, 16)
    rc = lib.fa_fwd_peers(null, null, null, null, null, 1, 1, 100, 128, s, s, s, s, 3, 1.0, 0, 0, peers, null, 0.0, 0,
                          ctypes.byref(am), null)
    assert rc < 0 and b"FP8" in lib.fa_last_error()
    rc = lib.fa_bwd_partial(null, null, null, null, null, null, null, null, null, null, 0, 1, 1, 100, 64, s, s, s, s, s, s, s,
                            1, 1.0, 0, 3, null, 0.0, 0, ctypes.byref(am), null)
    assert rc < 0 and b"cols" in lib.fa_last_error()


def test_attention_mask_packing_on_cpu():
    """AttentionMask: padded row pitch, transposed copy and 128 x 128 block summary (pure torch, no GPU needed)."""
    N = 300
    m = torch.zeros(2, 1, N, N, dtype=torch.bool)
    m[0, 0, 5, 290] = True        # block (0, 2)
    m[1, 0, 200, 130] = True      # block (1, 1)
    am = _native.AttentionMask(m)
    assert am.rows.shape == (2, 1, N, 48) and am.rows.stride(-2) == 48            # one bit per entry, 16 bytes per 128 keys
    assert am.rows[0, 0, 5, 290 >> 3] == 1 << (290 & 7) and am.cols[0, 0, 290, 5 >> 3] == 1 << 5
    assert am.rows[1, 0, 200, 130 >> 3] == 1 << (130 & 7) and (am.rows != 0).sum() == 2 and (am.cols != 0).sum() == 2
    want = torch.zeros(2, 1, 3, 3, dtype=torch.uint8)
    want[0, 0, 0, 2] = 1
    want[1, 0, 1, 1] = 1
    assert torch.equal(am.blocks, want)
    st = am.struct(2, 4, N, torch.device("cpu"))
    assert st.rows_strides[1] == 0 and st.rows_strides[2] == 48 and st.blocks_strides[2] == 3
    with pytest.raises(ValueError):
        am.struct(3, 4, N, torch.device("cpu"))
    assert _native.AttentionMask(torch.ones(N, N)).shape == (1, 1, N)
    win = _native.AttentionMask.sliding_window(300, 2, 1, device="cpu")       # band mask: no bytes, analytic block summary
    assert win.rows is None and win.window == (2, 1)
    i = torch.arange(300)
    dense = _native.AttentionMask(((i[None, :] - i[:, None]) >= -2) & ((i[None, :] - i[:, None]) <= 1))
    assert torch.equal(win.blocks > 0, dense.blocks > 0) and win.blocks[0, 0, 0, 2] == 0
    for left, right in ((0, 0), (127, 0), (128, 5), (300, 300), (1000, 0)):
        w = _native.AttentionMask.sliding_window(700, left, right, device="cpu")
        d = _native.AttentionMask(((i700 := torch.arange(700))[None, :] - i700[:, None] >= -left)
                                  & (i700[None, :] - i700[:, None] <= right))
        assert torch.equal(w.blocks > 0, d.blocks > 0), (left, right)
        assert ((w.blocks == 2) <= (d.blocks[..., :w.blocks.shape[-2], :] >= 1)).all()   # "full" only where something is visible
        inner = d.blocks[0, 0, :5, :5]   # blocks not touching the ragged edge: the two summaries agree exactly
        assert torch.equal(w.blocks[0, 0, :5, :5], inner), (left, right)
    st = win.struct(2, 4, 300, torch.device("cpu"))
    assert not st.rows and st.window_left == 2 and st.window_right == 1
    full = _native.AttentionMask(torch.ones(256, 256, dtype=torch.bool).tril())
    assert full.blocks.flatten().tolist() == [1, 0, 2, 1]   # diagonal blocks mixed, lower block fully visible, upper empty
    assert _native.AttentionMask(torch.ones(N, N)).blocks[0, 0].tolist() == [[2, 2, 1], [2, 2, 1], [1, 1, 1]]   # ragged edge


def test_cpp_autograd_node_loads_and_declines_what_it_does_not_serve():
    """csrc/torch_binding.cpp: the in-tree extension imports, binds the same ABI version as the ctypes handle, and refuses
    (-> Python Function, which owns the errors) everything but plain same-shape CUDA tensors at kernel head sizes."""
    node = fat._cpp_node()
    assert node is not None, "flash_attention_dlrs_b200/_fa_torch*.so is not built (python -c 'import __graft_entry__ as g; g.build()')"
    assert node.abi_version() == _lib.load().fa_version()
    q = torch.randn(1, 2, 128, 64)
    assert not node.supported(q, q, q)                      # CPU tensors
    with pytest.raises(NotImplementedError):
        fat.FlashAttention.apply(q, q, q)                   # ... raise like the reference, through the Python Function
    assert fat.FlashAttentionDeterministic.apply.__func__ is fat.FlashAttention.apply.__func__


def test_rectangular_entry_points_reject_what_they_do_not_take():
    """fa_fwd_rect / fa_bwd_rect (ABI v9): argument checks run before anything touches the GPU."""
    lib = _lib.load()
    s = (ctypes.c_int64 * 4)(64 * 128, 128 * 64, 64, 1)
    null = ctypes.c_void_p(0)
    # a key length of zero, float32 with Nk != Nq
    assert lib.fa_fwd_rect(null, null, null, null, null, 1, 1, 128, 0, 64, s, s, s, s, 1, 1.0, null) < 0
    assert b"Nk" in lib.fa_last_error()
    assert lib.fa_fwd_rect(null, null, null, null, null, 1, 1, 128, 256, 64, s, s, s, s, 2, 1.0, null) < 0
    assert b"float16 / bfloat16" in lib.fa_last_error()
    # null tensors are refused like in fa_fwd
    assert lib.fa_fwd_rect(null, null, null, null, null, 1, 1, 128, 256, 64, s, s, s, s, 1, 1.0, null) < 0
    assert b"null" in lib.fa_last_error()
    assert lib.fa_bwd_rect(null, null, null, null, null, null, null, null, null, 1, 1, 128, 256, 64, s, s, s, s, s, s, s, 1,
                           1.0, null) < 0
