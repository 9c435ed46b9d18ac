"""ctypes binding of libsatnerf_b200_dev.so (include/satnerf_b200_dev.h): the product library built with -DSNB_DEV_BUILD
(environment knobs SNB_TC_*, read once at load) plus microbenchmarks.  Used by profiles/*.py and one unit test only."""
import ctypes as C
import os

import torch

from . import capi

_DEV_PATH = os.environ.get("SNB_DEV_LIBRARY_PATH") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libsatnerf_b200_dev.so")
_SIGNATURES = {
    "snb_debug_read": (C.c_int, [C.c_void_p, C.c_size_t]),
    "snb_debug_hang_info": (C.c_int, [C.POINTER(C.c_uint)]),
    "snb_debug_mma_rate": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_longlong)]),
    "snb_debug_mma_ring2": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_longlong)]),
    "snb_debug_mma_ring": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_longlong)]),
    "snb_debug_dw_gemm": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "snb_last_error": (C.c_char_p, []),
}
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_DEV_PATH):
            raise RuntimeError(f"{_DEV_PATH} is missing: build it with `python -m satnerf_b200.build`")
        h = C.CDLL(_DEV_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(h, name)
            fn.restype, fn.argtypes = res, args
        _lib = h
    return _lib


def exported_symbols():
    return sorted(n for n in _SIGNATURES if n.startswith("snb_debug"))


def _check(code, what):
    if code != 0:
        raise RuntimeError(f"{what} failed ({code}): {lib().snb_last_error().decode()}")


def use_dev_library_for_product_calls():
    """Routes satnerf_b200.capi through the dev library (same entry points + SNB_TC_* knobs); call before the first capi.lib()."""
    capi._LIB_PATH = _DEV_PATH
    capi._lib = None


def debug_timestamps():
    """(64,4) int64 phase timestamps of the fused kernel's block 0 (probe builds)."""
    import numpy as np
    buf = np.zeros((64, 4), dtype=np.int64)
    _check(lib().snb_debug_read(buf.ctypes.data_as(C.c_void_p), buf.nbytes), "snb_debug_read")
    return buf


def debug_dw_gemm(xa, xb, k_splits=1):
    """out = xa^T xb on the tensor-core weight-gradient kernel."""
    P, Fa = xa.shape
    Fb = xb.shape[1]
    out = torch.empty(Fa, Fb, device=xa.device, dtype=torch.float32)
    tiles = (P + 127) // 128
    nbytes = tiles * (Fa // 64 + Fb // 64) * 16384 + ((Fa // 64 + 3) // 4) * ((Fb // 64 + 7) // 8) * k_splits * (256 * 512 * 4 + 128) + (1 << 16)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=xa.device)
    with torch.cuda.device(xa.device):
        _check(lib().snb_debug_dw_gemm(capi._ptr(xa), capi._ptr(xb), P, Fa, Fb, k_splits, capi._ptr(out), C.c_void_p(ws.data_ptr()), ws.numel(),
                                       capi._stream(xa.device)), "snb_debug_dw_gemm")
    return out
