"""Helpers for the -m gpu parity tests: call the C-ABI (include/aigv_b200.h) on torch-owned device buffers."""
import ctypes as C

import numpy as np
import torch

from aigv_assessor_b200 import _lib


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def stream():
    return C.c_void_p(int(torch.cuda.current_stream().cuda_stream))


def dev(x, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(x)) if isinstance(x, np.ndarray) else x
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda().contiguous()


def gemm(a_bf16, w_bf16, bias=None, colscale=None, residual=None, out_dtype=torch.bfloat16, gelu=False, out=None):
    lib = _lib.load()
    M, K = a_bf16.shape
    N = w_bf16.shape[0]
    if out is None:
        out = torch.empty((M, N), dtype=out_dtype, device="cuda")
    f32 = out.dtype == torch.float32
    st = lib.aigv_gemm_epilogue(ptr(a_bf16), ptr(w_bf16), ptr(bias), ptr(colscale), ptr(residual),
                                ptr(out) if f32 else None, None if f32 else ptr(out), M, N, K, int(gelu), stream())
    assert st == 0, lib.aigv_status_string(st)
    torch.cuda.synchronize()
    return out


def emax(y, ref):
    y = np.asarray(y, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    return float(np.abs(y - ref).max() / np.abs(ref).max())
