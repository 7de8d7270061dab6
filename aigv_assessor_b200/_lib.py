"""Build and load ``libaigv_b200.so`` (the C-ABI of include/aigv_b200.h) with ctypes.

The library is built IN-TREE next to this file (``nvcc -shared``, sm_100a only) so it travels with the repo snapshot;
there is no JIT cache and no CPU fallback: if the library is missing or does not load, every entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libaigv_b200.so")
SOURCES = ["api.cu", "gemm_tcgen05.cu", "attention2_tcgen05.cu", "rowops.cu", "preprocess.cu",
           "llm_glue.cu", "llm_prefill.cu", "motion_conv.cu"]
HEADERS = ["kernels.h", "ptx.cuh", os.path.join("..", "..", "include", "aigv_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-t", "0",
              "-Xcompiler", "-fPIC", "-shared"]


class AigvError(RuntimeError):
    pass


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise AigvError("nvcc not found; cannot build libaigv_b200.so")


def source_hash() -> str:
    """sha256 over every source / header the library is built from (compiled in as AIGV_SRC_HASH, returned by
    aigv_source_hash()): load() compares it with the sources on disk, so an edited .cu can never silently run an old
    binary, wherever file times got lost (the snapshot that travels to the GPU box)."""
    import hashlib
    h = hashlib.sha256()
    for s in sorted(SOURCES + HEADERS):
        with open(os.path.join(CSRC, s), "rb") as f:
            h.update(s.encode())
            h.update(f.read())
    return h.hexdigest()[:16]


def _built_hash(path: str) -> str | None:
    # read from the file's bytes, not through dlopen: a library dlopen()ed before a rebuild stays mapped under the
    # same path, and the rebuilt file would then look stale to this process
    try:
        with open(path, "rb") as f:
            data = f.read()
    except OSError:
        return None
    i = data.find(b"aigv-src-hash:")
    if i < 0:
        return None
    j = data.find(b"\0", i)
    return data[i + 14:j].decode(errors="replace")


def _stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    return _built_hash(LIB_PATH) != source_hash()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a into libaigv_b200.so (cross-compiles without a GPU)."""
    if not force and not _stale():
        return LIB_PATH
    extra = os.environ.get("AIGV_NVCC_EXTRA", "").split()  # experiments only, e.g. -DAIGV_ATTN_MIN_CTAS=3
    cmd = [_nvcc(), *NVCC_FLAGS, *extra, f'-DAIGV_SRC_HASH="{source_hash()}"', "-o", LIB_PATH,
           *[os.path.join(CSRC, s) for s in SOURCES]]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise AigvError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB_PATH


class Cfg(C.Structure):
    _fields_ = [("hidden_size", C.c_int32), ("num_heads", C.c_int32), ("intermediate_size", C.c_int32),
                ("num_layers", C.c_int32), ("image_size", C.c_int32), ("patch_size", C.c_int32),
                ("llm_hidden", C.c_int32), ("qkv_bias", C.c_int32), ("qk_normalization", C.c_int32),
                ("norm_type", C.c_int32), ("layer_norm_eps", C.c_float), ("mlp1_ln_eps", C.c_float),
                ("max_tiles", C.c_int32)]


class LlmCfg(C.Structure):
    _fields_ = [("hidden_size", C.c_int32), ("num_heads", C.c_int32), ("num_kv_heads", C.c_int32),
                ("intermediate_size", C.c_int32), ("num_layers", C.c_int32), ("vocab_size", C.c_int32),
                ("rms_norm_eps", C.c_float), ("rope_theta", C.c_float), ("max_tokens", C.c_int32),
                ("max_positions", C.c_int32)]


# name -> (restype, argtypes): exactly the symbols include/aigv_b200.h declares
_P = C.c_void_p
_F = C.POINTER(C.c_float)
SIGNATURES = {
    "aigv_version": (C.c_char_p, []),
    "aigv_source_hash": (C.c_char_p, []),
    "aigv_status_string": (C.c_char_p, [C.c_int]),
    "aigv_create": (C.c_int, [C.POINTER(Cfg), C.POINTER(_P)]),
    "aigv_destroy": (None, [_P]),
    "aigv_last_error": (C.c_char_p, [_P]),
    "aigv_workspace_bytes": (C.c_size_t, [_P, C.c_int]),
    "aigv_set_weight": (C.c_int, [_P, C.c_char_p, _P, C.c_int, C.POINTER(C.c_int64), C.c_int]),
    "aigv_finalize_weights": (C.c_int, [_P, _P]),
    "aigv_extract_feature": (C.c_int, [_P, _P, C.c_int, _P, _P]),
    "aigv_copy_vit_hidden": (C.c_int, [_P, _P, C.c_int, _P]),
    "aigv_profile_enable": (C.c_int, [_P, C.c_int]),
    "aigv_profile_num_classes": (C.c_int, []),
    "aigv_profile_class_name": (C.c_char_p, [C.c_int]),
    "aigv_profile_read": (C.c_int, [_P, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.c_int, C.c_int]),
    "aigv_launch_count": (C.c_int64, [_P]),
    "aigv_tile_grid": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "aigv_preprocess_u8": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int, C.c_int,
                                     C.c_int, _P, _P, _P]),
    "aigv_gemm_epilogue": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "aigv_stats_slices": (C.c_int, [C.c_int]),
    "aigv_fold_norm": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, C.c_int, C.c_int, _P]),
    "aigv_gemm_residual_stats": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, _P]),
    "aigv_gemm_norm_fused": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_float, _P, C.c_int, C.c_int,
                                       C.c_int, C.c_int, _P]),
    "aigv_assemble_tokens_stats": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "aigv_attention": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, _P]),
    "aigv_exp2_poly": (C.c_int, [_P, _P, C.c_int, _P]),
    "aigv_gemm_tile_order": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int]),
    "aigv_splice_tokens": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, _P]),
    "aigv_mlpscore": (C.c_int, [_P, C.c_int, C.POINTER(_P), C.POINTER(_P), _P, _P, C.c_int, _P]),
    "aigv_llm_create": (C.c_int, [C.POINTER(LlmCfg), C.POINTER(_P)]),
    "aigv_llm_destroy": (None, [_P]),
    "aigv_llm_last_error": (C.c_char_p, [_P]),
    "aigv_llm_launch_count": (C.c_int64, [_P]),
    "aigv_llm_set_weight": (C.c_int, [_P, C.c_char_p, _P, C.c_int, C.POINTER(C.c_int64), C.c_int]),
    "aigv_llm_finalize_weights": (C.c_int, [_P, _P]),
    "aigv_llm_prefill": (C.c_int, [_P, _P, C.c_int, C.c_int, _P]),
    "aigv_llm_hidden_rows": (C.c_int, [_P, C.POINTER(C.c_int32), C.c_int, _P, _P]),
    "aigv_llm_logits_rows": (C.c_int, [_P, C.POINTER(C.c_int32), C.c_int, _P, _P]),
    "aigv_llm_copy_residual": (C.c_int, [_P, _P, _P]),
    "aigv_rope": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, _P]),
    "aigv_attention_gqa": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, _P]),
    "aigv_im2col3d": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_int64,
                                C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "aigv_gemm_ex": (C.c_int, [_P, C.c_int, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "aigv_maxpool3d": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                 C.c_int, _P]),
    "aigv_add_relu": (C.c_int, [_P, _P, _P, C.c_int64, C.c_int, C.c_int, _P]),
    "aigv_slowfast_tail": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.c_int, C.c_int, C.c_int, _P]),
    "aigv_layernorm": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.c_float, C.c_int, _P]),
    "aigv_rmsnorm_inplace": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, _P]),
    "aigv_im2col": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "aigv_assemble_tokens": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.c_int, _P]),
    "aigv_pixel_shuffle": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, _P]),
}

_lib = None


def load(build_if_missing: bool = True):
    """dlopen the library and bind every declared symbol. Raises AigvError when it cannot (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("AIGV_LIB")  # development: A/B a prebuilt variant of the same ABI (no staleness check)
    if not path:
        path = LIB_PATH
        if _stale():  # missing, or built from other sources than the ones on disk
            if not build_if_missing:
                raise AigvError(f"{LIB_PATH} is missing or stale; run `python -c 'import __graft_entry__ as g; g.build()'`")
            build(force=True)
    try:
        lib = C.CDLL(path)
    except OSError as e:  # pragma: no cover
        raise AigvError(f"cannot load {path}: {e}") from e
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise AigvError(f"{LIB_PATH} does not export {name}") from e
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int, handle=None, what: str = ""):
    if status == 0:
        return
    lib = load()
    msg = lib.aigv_status_string(status).decode()
    if handle:
        detail = lib.aigv_last_error(handle)
        if detail:
            msg += ": " + detail.decode()
    if status in (-1, -4, -5):
        raise ValueError(f"{what}: {msg}" if what else msg)
    raise AigvError(f"{what}: {msg}" if what else msg)
