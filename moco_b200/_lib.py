"""ctypes binding of libmoco_b200.so (the C ABI in include/moco_b200.h).

The library is built in-tree by ``moco_b200/build.py`` (nvcc, sm_100a).  There is
no CPU fallback: if the shared object is missing and cannot be built, importing
any compute entry point raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int64, c_size_t, c_uint32, c_void_p

from . import build as _build

ABI_VERSION = 3                    # MOCO_B200_ABI_VERSION (include/moco_b200.h)
MOCO_F32, MOCO_BF16 = 0, 1
NCE_AUTO, NCE_FORCE_SIMT, NCE_CTA_PAIR, NCE_SINGLE_CTA = 0, 1, 2, 4
NCE_TWO_PASS, NCE_ONE_PASS = 512, 1024
ONE_PASS_MAX_INV_T = 25.0          # MOCO_ONE_PASS_MAX_INV_T (include/moco_b200.h)
GATHER_AUTO, GATHER_LDG = 0, 1

# every symbol include/moco_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "moco_abi_version": (c_int, []),
    "moco_last_error": (c_char_p, []),
    "moco_device_info": (c_int, [POINTER(c_int), POINTER(c_int), POINTER(c_int)]),
    "moco_nce_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "moco_nce_fwd": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_float,
                             c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                             c_void_p, c_size_t, c_int, c_void_p]),
    "moco_nce_step": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_float,
                              c_void_p, c_int, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                              c_void_p, c_size_t, c_int, c_void_p]),
    "moco_prof_sweep_window": (c_int, [c_void_p, c_int, POINTER(c_float), c_void_p]),
    "moco_prof_set_events": (c_int, [c_int, c_void_p, c_void_p]),
    "moco_nce_bwd_dense": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_float,
                                   c_void_p, c_void_p]),
    "moco_queue_enqueue": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int64, c_int64, c_void_p]),
    "moco_nce_shard_stats": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_float, c_void_p,
                                     c_void_p, c_size_t, c_int, c_void_p]),
    "moco_nce_shard_merge": (c_int, [c_void_p, c_int, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_void_p, c_size_t, c_void_p]),
    "moco_nce_shard_dq": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_void_p,
                                  c_void_p, c_size_t, c_int, c_void_p]),
    "moco_nce_shard_dq_finish": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_float, c_void_p, c_void_p]),
    "moco_nce_shard_dq_finish_peers": (c_int, [POINTER(c_void_p), c_int, c_int, c_void_p, c_int, c_void_p, c_int, c_int,
                                               c_float, c_void_p, c_void_p]),
    "moco_queue_enqueue_shard": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int64, c_int64,
                                         c_int64, c_int64, c_void_p]),
    "moco_f32_to_bf16": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "moco_ema_chunk_elems": (c_int, []),
    "moco_ema_update": (c_int, [c_void_p, c_void_p, c_int, c_int, c_float, c_float, c_void_p]),
    "moco_crop_s2d_bf16": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "moco_maxpool3x3s2_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "moco_maxpool3x3s2_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "moco_bn_workspace_bytes": (c_size_t, []),
    "moco_bn_fwd_train": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_float, c_float, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "moco_bn_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                            c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "moco_crop_to_nhwc_bf16": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_int, c_int, c_int, c_void_p]),
    "moco_shuffle_gather": (c_int, [POINTER(c_void_p), c_int, c_int, c_void_p, c_int, c_size_t, c_void_p, c_int, c_void_p]),
    "moco_shuffle_gather_sync": (c_int, [POINTER(c_void_p), POINTER(c_void_p), c_int, c_int, c_uint32, c_int, c_void_p, c_int,
                                         c_size_t, c_void_p, c_int, c_void_p]),
    "moco_p2p_last_timeout": (c_int, [POINTER(c_uint32)]),
    "moco_crop_gather_nhwc_bf16": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "moco_signal_barrier": (c_int, [POINTER(c_void_p), c_int, c_int, c_uint32, c_void_p]),
    "moco_p2p_alloc": (c_int, [c_size_t, POINTER(c_void_p), c_void_p]),
    "moco_p2p_open": (c_int, [c_void_p, POINTER(c_void_p)]),
    "moco_p2p_close": (c_int, [c_void_p]),
    "moco_p2p_free": (c_int, [c_void_p]),
}

_lib = None


def lib_path() -> str:
    return _build.LIB


def load() -> ctypes.CDLL:
    """Load (building first if needed and possible) libmoco_b200.so."""
    global _lib
    if _lib is not None:
        return _lib
    import torch  # noqa: F401  (loads the CUDA runtime the library links against)
    path = _build.LIB
    if not os.path.exists(path):
        try:
            _build.build()
        except Exception as exc:  # no nvcc on this box and no prebuilt library
            raise RuntimeError(
                f"moco_b200: {path} is missing and could not be built ({exc}); "
                "there is no CPU fallback for the CUDA hot path") from exc
    lib = ctypes.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.moco_abi_version() != ABI_VERSION:
        raise RuntimeError("moco_b200: ABI version mismatch between _lib.py and libmoco_b200.so")
    _lib = _Counting(lib)
    return _lib


# kernels launched per successful C-ABI call (bench.py reports the total as `gpu_launches`)
launches = 0


class _Counting:
    """Thin proxy over the CDLL that counts this library's kernel launches."""

    _PER_CALL = {"moco_nce_shard_stats": 3, "moco_nce_shard_merge": 1, "moco_nce_shard_dq": 2,
                 "moco_nce_shard_dq_finish": 1, "moco_nce_shard_dq_finish_peers": 1, "moco_queue_enqueue_shard": 1, "moco_queue_enqueue": 1, "moco_f32_to_bf16": 1, "moco_shuffle_gather": 1, "moco_shuffle_gather_sync": 1, "moco_crop_gather_nhwc_bf16": 1,
                 "moco_ema_update": 1, "moco_crop_to_nhwc_bf16": 1, "moco_bn_fwd_train": 2, "moco_bn_bwd": 2, "moco_crop_s2d_bf16": 1, "moco_maxpool3x3s2_fwd": 1, "moco_maxpool3x3s2_bwd": 1,
                 "moco_signal_barrier": 1, "moco_nce_bwd_dense": 1}

    def __init__(self, lib):
        self._lib = lib
        for name in SIGNATURES:
            fn = getattr(lib, name)
            if name == "moco_nce_fwd":
                setattr(self, name, self._wrap_nce(fn))
            elif name == "moco_nce_step":
                setattr(self, name, self._wrap_step(fn))
            elif name == "moco_nce_shard_dq":      # one-pass finish: dq_reduce only; two-pass: dq kernel + dq_reduce
                setattr(self, name, self._wrap_flags(fn, 11))
            elif name in self._PER_CALL:
                setattr(self, name, self._wrap(fn, self._PER_CALL[name]))
            else:
                setattr(self, name, fn)

    @staticmethod
    def _wrap(fn, n):
        def call(*a):
            global launches
            rc = fn(*a)
            if rc == 0:
                launches += n
            return rc
        return call

    @staticmethod
    def _wrap_flags(fn, flag_index):
        def call(*a):
            global launches
            rc = fn(*a)
            if rc == 0:
                launches += 1 if (a[flag_index] & NCE_ONE_PASS) else 2
            return rc
        return call

    @staticmethod
    def _head_launches(C, inv_T, flags, dq, logits, f32):
        """Kernels one head evaluation launches (mirrors the dispatch in csrc/capi.cu)."""
        simt = bool(flags & NCE_FORCE_SIMT) or C % 64 != 0 or C > 256
        if simt:
            return 2                                             # prep + row kernel
        one_pass = (dq and not logits and not (flags & NCE_TWO_PASS)
                    and ((flags & NCE_ONE_PASS) or inv_T <= ONE_PASS_MAX_INV_T))
        if one_pass:                                             # sweep + tail (+ prep for the bf16 copy at C > 128)
            return 2 + (1 if (C > 128 and f32) else 0)
        return 5 if dq else 3                                    # prep + stats + combine [+ dq + dq_reduce]

    @classmethod
    def _wrap_nce(cls, fn):
        def call(*a):
            global launches
            rc = fn(*a)
            if rc == 0:
                launches += cls._head_launches(a[5], a[7], a[16], a[13], a[8], a[2] == MOCO_F32)
            return rc
        return call

    @classmethod
    def _wrap_step(cls, fn):
        def call(*a):
            global launches
            rc = fn(*a)
            if rc == 0:
                n = cls._head_launches(a[7], a[9], a[22], True, False, a[2] == MOCO_F32)
                fused = n <= 3 and a[7] % 8 == 0 and 256 % (a[7] // 8) == 0      # the tail kernel also enqueues
                launches += n + (0 if (fused or a[12] == 0) else 1)
            return rc
        return call


def check(code: int, what: str) -> None:
    if code != 0:
        msg = load().moco_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"moco_b200.{what} failed ({code}): {msg}")


def dtype_code(t) -> int:
    import torch
    if t.dtype == torch.float32:
        return MOCO_F32
    if t.dtype == torch.bfloat16:
        return MOCO_BF16
    raise TypeError(f"moco_b200: unsupported dtype {t.dtype} (expected float32 or bfloat16)")


def cur_stream() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream


def require_cuda(*tensors) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("moco_b200: the contrastive hot path runs on CUDA only "
                               "(got a CPU tensor); there is no CPU fallback")
