"""Loader (and in-tree builder) of libfa_b200.so, the C-ABI library declared in include/fa_b200.h.

The product path has no CPU or eager fallback: if the shared library is missing or a call fails, the
error is raised to the caller (`FlashAttentionLibraryError`).
"""
from __future__ import annotations

import ctypes
import os
import shutil
import subprocess
import threading
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
REPO_ROOT = PKG_DIR.parent
CSRC = PKG_DIR / "csrc"
# FA_B200_LIB points at an alternative build of the same ABI (profiling / ablation builds); default is in-tree.
LIB_PATH = Path(os.environ["FA_B200_LIB"]) if os.environ.get("FA_B200_LIB") else PKG_DIR / "libfa_b200.so"
HEADER = REPO_ROOT / "include" / "fa_b200.h"

FA_DTYPE_F16, FA_DTYPE_BF16, FA_DTYPE_F32, FA_DTYPE_F8E4M3, FA_DTYPE_F8E5M2 = 0, 1, 2, 3, 4

# every symbol include/fa_b200.h declares
EXPORTED_SYMBOLS = (
    "fa_version",
    "fa_last_error",
    "fa_fwd",
    "fa_fwd_peers",
    "fa_bwd_preprocess",
    "fa_bwd_workspace_bytes",
    "fa_bwd",
    "fa_bwd_partial",
    "fa_fwd_rect",
    "fa_bwd_rect",
    "fa_merge_partial",
    "fa_accumulate",
    "fa_round_rows",
)


class FlashAttentionLibraryError(RuntimeError):
    """libfa_b200.so is missing, failed to load, or one of its entry points returned an error."""


# -lineinfo only adds line tables for `ncu --import-source`: the SASS is the same with and without it
# (profiles/r02_build_flags.txt compares the two instruction streams).  FA_B200_DEBUG=1 builds the debug flavour:
# mbarrier watchdog on (sm100_ptx.cuh).
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
] + (["-DFA_WATCHDOG=1"] if os.environ.get("FA_B200_DEBUG") == "1" else [])
OBJ_DIR = REPO_ROOT / "build" / "obj"


def _translation_units():
    """csrc/*.cu: the C ABI (fa_api.cu) and one file per tcgen05 kernel family, compiled in parallel."""
    return sorted(CSRC.glob("*.cu"))


def _sources():
    return _translation_units() + sorted(CSRC.glob("*.cuh")) + sorted(CSRC.glob("*.h")) + [HEADER]


def needs_build() -> bool:
    if not LIB_PATH.exists():
        return True
    t = LIB_PATH.stat().st_mtime
    return any(s.stat().st_mtime > t for s in _sources())


def build(force: bool = False, verbose: bool = False, extra_flags=()) -> Path:
    """Compile csrc/*.cu for sm_100a (one nvcc process per translation unit, in parallel) and link libfa_b200.so next
    to this file.  nvcc cross-compiles without a GPU."""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise FlashAttentionLibraryError("nvcc not found; cannot build libfa_b200.so")
    OBJ_DIR.mkdir(parents=True, exist_ok=True)
    tag = "%d" % os.getpid()
    flags = [*NVCC_FLAGS, *extra_flags] + (["-Xptxas", "-v"] if verbose else [])
    jobs = []
    for src in _translation_units():
        obj = OBJ_DIR / f"{src.stem}.{tag}.o"
        jobs.append((src, obj, subprocess.Popen([nvcc, *flags, "-c", "-o", str(obj), str(src)],
                                                stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    logs, failed = [], []
    for src, obj, proc in jobs:
        out, _ = proc.communicate()
        logs.append(out)
        if proc.returncode != 0:
            failed.append(f"{src.name}:\n{out}")
    try:
        if failed:
            raise FlashAttentionLibraryError("nvcc failed:\n" + "\n".join(failed))
        if verbose:
            print("".join(logs))
        tmp = LIB_PATH.with_suffix(".so.tmp" + tag)
        link = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", str(tmp),
                               *[str(obj) for _, obj, _ in jobs]], capture_output=True, text=True)
        if link.returncode != 0:
            raise FlashAttentionLibraryError("link failed:\n" + link.stdout + link.stderr)
        os.replace(tmp, LIB_PATH)
    finally:
        for _, obj, _ in jobs:
            if obj.exists():
                obj.unlink()
    return LIB_PATH


# ---------------------------------------------------------------------------------------------------------------
# C++ autograd node over the C ABI (csrc/torch_binding.cpp): the plain FlashAttention.apply(Q, K, V[, causal, scale])
# without the interpreter on the launch path.  Host plumbing only — it calls the same entry points of libfa_b200.so.
TORCH_BINDING_SRC = CSRC / "torch_binding.cpp"


def torch_binding_path() -> Path:
    import sysconfig
    return PKG_DIR / ("_fa_torch" + sysconfig.get_config_var("EXT_SUFFIX"))


def torch_binding_needs_build() -> bool:
    out = torch_binding_path()
    if not out.exists():
        return True
    t = out.stat().st_mtime
    return TORCH_BINDING_SRC.stat().st_mtime > t or HEADER.stat().st_mtime > t


def build_torch_binding(force: bool = False) -> Path:
    """g++ csrc/torch_binding.cpp -> _fa_torch<EXT_SUFFIX> next to libfa_b200.so (which it links by $ORIGIN)."""
    out = torch_binding_path()
    if not force and not torch_binding_needs_build():
        return out
    import sysconfig
    import torch
    from torch.utils import cpp_extension
    cxx = shutil.which("g++") or "g++"
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    cmd = [cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-DTORCH_EXTENSION_NAME=_fa_torch",
           "-DTORCH_API_INCLUDE_EXTENSION_H", "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI),
           "-I", str(REPO_ROOT / "include"), "-I", sysconfig.get_paths()["include"], "-I", cuda_inc]
    for inc in cpp_extension.include_paths():
        cmd += ["-isystem", inc]
    tmp = out.with_suffix(".tmp%d" % os.getpid())
    cmd += [str(TORCH_BINDING_SRC), "-o", str(tmp), "-L", str(PKG_DIR), "-l:libfa_b200.so", "-Wl,-rpath,$ORIGIN"]
    for lp in cpp_extension.library_paths():
        cmd += ["-L", lp]
    cmd += ["-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch", "-ltorch_python"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        if tmp.exists():
            tmp.unlink()
        raise FlashAttentionLibraryError("building the torch binding failed:\n" + r.stdout + r.stderr)
    os.replace(tmp, out)
    return out


_lock = threading.Lock()
_lib = None

_I64x4 = ctypes.c_int64 * 4
_I64x3 = ctypes.c_int64 * 3


class AttnMaskStruct(ctypes.Structure):
    """fa_attn_mask of include/fa_b200.h."""
    _fields_ = [("rows", ctypes.c_void_p), ("rows_strides", _I64x3),
                ("cols", ctypes.c_void_p), ("cols_strides", _I64x3),
                ("blocks", ctypes.c_void_p), ("blocks_strides", _I64x3),
                ("window_left", ctypes.c_int32), ("window_right", ctypes.c_int32)]


def _declare(lib):
    vp, i, f, sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_size_t
    st = ctypes.POINTER(ctypes.c_int64)
    lib.fa_version.restype = i
    lib.fa_version.argtypes = []
    lib.fa_last_error.restype = ctypes.c_char_p
    lib.fa_last_error.argtypes = []
    lib.fa_fwd.restype = i
    lib.fa_fwd.argtypes = [vp, vp, vp, vp, vp, i, i, i, i, st, st, st, st, i, f, i, vp]
    lib.fa_fwd_peers.restype = i
    lib.fa_fwd_peers.argtypes = [vp, vp, vp, vp, vp, i, i, i, i, st, st, st, st, i, f, i, i, ctypes.POINTER(vp), vp,
                                 f, ctypes.c_uint64, ctypes.POINTER(AttnMaskStruct), vp]
    lib.fa_bwd_preprocess.restype = i
    lib.fa_bwd_preprocess.argtypes = [vp, vp, vp, i, i, i, i, st, st, i, vp]
    lib.fa_bwd_workspace_bytes.restype = sz
    lib.fa_bwd_workspace_bytes.argtypes = [i, i, i, i, i, i, i]
    lib.fa_bwd.restype = i
    lib.fa_bwd.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, sz, i, i, i, i, st, st, st, st, st, st, st, i, f, i, vp]
    lib.fa_bwd_partial.restype = i
    lib.fa_bwd_partial.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, sz, i, i, i, i, st, st, st, st, st, st, st, i, f, i,
                                   i, vp, f, ctypes.c_uint64, ctypes.POINTER(AttnMaskStruct), vp]


def _declare_rect(lib):
    vp, i, f = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
    st = ctypes.POINTER(ctypes.c_int64)
    lib.fa_fwd_rect.restype = i
    lib.fa_fwd_rect.argtypes = [vp, vp, vp, vp, vp, i, i, i, i, i, st, st, st, st, i, f, vp]
    lib.fa_bwd_rect.restype = i
    lib.fa_bwd_rect.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, i, i, i, i, i, st, st, st, st, st, st, st, i, f, vp]


def _declare_ring(lib):
    vp, i, ll = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong
    lib.fa_merge_partial.restype = i
    lib.fa_merge_partial.argtypes = [vp, vp, vp, vp, ll, i, i, i, vp]
    lib.fa_accumulate.restype = i
    lib.fa_accumulate.argtypes = [vp, vp, ll, i, i, vp]
    lib.fa_round_rows.restype = i
    lib.fa_round_rows.argtypes = [vp, vp, ll, i, vp]


def load():
    """Return the ctypes handle of libfa_b200.so (loaded once per process).  Never builds implicitly."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not LIB_PATH.exists():
                raise FlashAttentionLibraryError(
                    f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                    "(nvcc, sm_100a). There is no CPU or eager fallback for the attention path."
                )
            try:
                lib = ctypes.CDLL(str(LIB_PATH))
            except OSError as e:  # pragma: no cover
                raise FlashAttentionLibraryError(f"cannot load {LIB_PATH}: {e}") from e
            for sym in EXPORTED_SYMBOLS:
                if not hasattr(lib, sym):
                    raise FlashAttentionLibraryError(f"{LIB_PATH} does not export {sym}")
            _declare(lib)
            _declare_rect(lib)
            _declare_ring(lib)
            _lib = lib
    return _lib


def strides4(t) -> "ctypes.Array":
    return _I64x4(*t.stride())


def check(rc: int, what: str):
    if rc != 0:
        msg = load().fa_last_error().decode("utf-8", "replace")
        raise FlashAttentionLibraryError(f"{what} failed (code {rc}): {msg}")
