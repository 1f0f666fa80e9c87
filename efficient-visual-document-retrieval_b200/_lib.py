"""ctypes binding of libevdr.so (include/evdr.h).  There is no CPU fallback: if the
shared library is missing, or no sm_100 device is present, the first compute call raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from pathlib import Path

import torch

_PKG = Path(__file__).resolve().parent
_LIB_PATH = _PKG / "libevdr.so"

EVDR_F32, EVDR_BF16 = 0, 1
CHUNK = 16
TC_DIM = 128
TC_LQ = 32
TC_LQ_MAX = 128      # longer queries are scored as up to 4 groups of 32 tokens
TOPK_MAX = 128


def row_blocks(k_blocks: int) -> int:
    """128-wide K blocks STORED per operand row: 1 for plain bf16 (k_blocks = 1), 2 = [hi | lo] for the bf16x3
    split (k_blocks = 3: the kernel forms the three products hi.hi + hi.lo + lo.hi from the two blocks)."""
    return 2 if k_blocks == 3 else 1

_lib = None
_lock = threading.RLock()
_handles: dict[int, C.c_void_p] = {}

_vp, _i, _i64 = C.c_void_p, C.c_int, C.c_int64

_SIGNATURES = {
    "evdr_version": ([], _i),
    "evdr_last_error": ([], C.c_char_p),
    "evdr_last_kernel": ([], C.c_char_p),
    "evdr_create": ([_i, C.POINTER(_vp)], _i),
    "evdr_destroy": ([_vp], _i),
    "evdr_sm_count": ([_vp], _i),
    "evdr_maxsim_fwd": ([_vp, _vp, _i, _vp, _i, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp], _i),
    "evdr_maxsim_bwd": ([_vp, _vp, _vp, _i, _vp, _i, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp], _i),
    "evdr_pack_plan": ([_vp, _vp, _i, _i, _i, _vp, _vp], _i),
    "evdr_pack_pages": ([_vp, _vp, _i, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp], _i),
    "evdr_pack_plan_ragged": ([_vp, _vp, _vp, _i, _vp, _vp], _i),
    "evdr_pack_pages_ragged": ([_vp, _vp, _i, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp], _i),
    "evdr_pack_queries": ([_vp, _vp, _i, _vp, _i, _i, _i, _i, _vp, _vp, _vp], _i),
    "evdr_maxsim_packed_fwd": ([_vp, _vp, _vp, _i, _vp, _i64, _vp, _i, _i64, _i, _i, _vp, _i64, _vp], _i),
    "evdr_maxsim_packed_argmax": ([_vp, _vp, _vp, _i, _i, _vp, _i64, _vp, _i, _i64, _i, _vp, _vp], _i),
    "evdr_maxsim_rescore": ([_vp, _vp, _i, _vp, _i, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp], _i),
    "evdr_maxsim_train_fwd": ([_vp, _vp, _i, _vp, _i, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp], _i),
    "evdr_maxsim_packed_argmax_scores": ([_vp, _vp, _vp, _i, _i, _vp, _i64, _vp, _i, _i64, _i, _vp, _vp, _i64, _vp], _i),
    "evdr_slots_to_tokens": ([_vp, _vp, _vp, _i, _i, _i, _vp], _i),
    "evdr_group_sum": ([_vp, _vp, _i64, _i, _i, _i, _vp, _i64, _vp], _i),
    "evdr_hard_token_argmax": ([_vp, _vp, _i, _vp, _i, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp], _i),
    "evdr_topk": ([_vp, _vp, _i64, _i, _i, _i, _vp, _i64, _vp, _vp, _vp], _i),
    "evdr_topk_merge": ([_vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp], _i),
    "evdr_retrieval_metrics": ([_vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _i, _vp, _vp], _i),
    "evdr_topk_metrics": ([_vp, _vp, _i64, _i, _i, _i, _vp, _i64, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp], _i),
    "evdr_comm_unique_id": ([_vp], _i),
    "evdr_comm_init": ([_vp, _vp, _i, _i], _i),
    "evdr_comm_world": ([_vp], _i),
    "evdr_sharded_topk": ([_vp, _vp, _i64, _i, _i, _i, _i64, _vp, _vp, _vp], _i),
    "evdr_allgather_score_rows": ([_vp, _vp, _i64, _i, _i, _vp, _i64, _vp], _i),
    "evdr_peer_region_init": ([_vp, _i, _i, _i, _vp], _i),
    "evdr_peer_connect": ([_vp, _vp], _i),
    "evdr_peer_world": ([_vp], _i),
    "evdr_peer_sharded_topk": ([_vp, _vp, _i64, _i, _i, _i, _i64, _vp, _vp, _vp], _i),
    "evdr_sharded_search": ([_vp, _vp, _vp, _i, _vp, _i64, _vp, _i, _i64, _i, _i64, _i, _vp, _vp, _vp], _i),
    "evdr_l2norm_fwd": ([_vp, _vp, _vp, _i64, _i, _vp, _vp], _i),
    "evdr_l2norm_bwd": ([_vp, _vp, _vp, _vp, _i64, _i, _vp, _vp], _i),
    "evdr_rank_gap": ([_vp, _vp, _i64, _vp, _i64, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp], _i),
    "evdr_liscore_loss": ([_vp, _vp, _i64, _vp, _i64, _i, _i, _i, C.c_float, _vp, _vp, _vp, _vp, _vp], _i),
    "evdr_pairwise_loss": ([_vp, _vp, _i64, _vp, _i64, _i, _i, _vp, _vp, _vp, _vp], _i),
    "evdr_lambda_loss": ([_vp, _vp, _i64, _vp, _i64, _i, _i, C.c_float, C.c_float, _vp, _vp, _vp, _vp], _i),
    "evdr_ranknce_loss": ([_vp, _vp, _i64, _vp, _i64, _i, _i, C.c_float, C.c_float, _vp, _vp, _vp, _vp], _i),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


class EvdrError(RuntimeError):
    pass


def lib_path() -> Path:
    return _LIB_PATH


def load() -> C.CDLL:
    """Load libevdr.so (built in-tree by csrc/build.py).  Raises if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not _LIB_PATH.exists():
            raise EvdrError(
                f"{_LIB_PATH} not found: build it with `python __graft_entry__.py` or "
                f"`python {_PKG.name}/csrc/build.py`.  There is no CPU or PyTorch fallback."
            )
        lib = C.CDLL(str(_LIB_PATH), mode=os.RTLD_NOW if hasattr(os, "RTLD_NOW") else C.DEFAULT_MODE)
        for name, (argtypes, restype) in _SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the symbol is not exported
            fn.argtypes = argtypes
            fn.restype = restype
        _lib = lib
    return _lib


def last_error() -> str:
    msg = load().evdr_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def last_kernel() -> str:
    """Name + template arguments of the scoring kernel this thread launched last (evdr_last_kernel)."""
    msg = load().evdr_last_kernel()
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc: int, what: str = "") -> None:
    if rc == 0:
        return
    msg = last_error()
    if rc == -1:
        raise ValueError(f"{what}: {msg}")
    if rc == -3:
        raise NotImplementedError(f"{what}: {msg}")
    raise EvdrError(f"{what}: {msg} (status {rc})")


def handle(device: torch.device | int) -> C.c_void_p:
    """Per-device evdr handle (created lazily, kept for the life of the process)."""
    idx = device if isinstance(device, int) else (device.index if device.index is not None else torch.cuda.current_device())
    h = _handles.get(idx)
    if h is None:
        with _lock:                      # two threads must not create two handles for one device
            h = _handles.get(idx)
            if h is None:
                lib = load()
                out = _vp()
                check(lib.evdr_create(idx, C.byref(out)), "evdr_create")
                _handles[idx] = h = out
    return h


def require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise EvdrError(
            f"{name} is on {t.device}: this package runs on CUDA (B200, sm_100a) only and has no CPU fallback"
        )


def ptr(t: torch.Tensor | None) -> int | None:
    return None if t is None else t.data_ptr()


def stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return EVDR_F32
    if t.dtype == torch.bfloat16:
        return EVDR_BF16
    raise TypeError(f"unsupported dtype {t.dtype}")
