"""ctypes binding of ``libkzd.so`` (the C ABI declared in ``include/kzd.h``).

There is no fallback: if the shared library is missing or a call fails this module raises.
torch is used by the callers only for device memory and streams; nothing here takes a torch type.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("KZD_LIB_PATH") or os.path.join(_HERE, "libkzd.so")  # override: A/B builds of the same ABI

KZD_MAX_LEVELS = 8
KZD_MAX_TOPK = 32
KZD_ASSIGN_PERSISTENT = 1        # include/kzd.h: flags of kzd_tal_assign
KZD_ASSIGN_TRUSTED_ANCHORS = 2
KZD_ASSIGN_NATIVE_SCORES = 4
KZD_ASSIGN_WARP_PER_GT = 8


class KzdError(RuntimeError):
    pass


class Level(C.Structure):
    """``kzd_level``: one pyramid level of the anchor grid."""

    _fields_ = [("h", C.c_int32), ("w", C.c_int32), ("offset", C.c_int32), ("stride", C.c_float)]


def _load():
    if not os.path.exists(LIB_PATH):
        raise KzdError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C kuzushiji_vision_b200/csrc`). There is no CPU or torch fallback."
        )
    lib = C.CDLL(LIB_PATH)
    vp, i32, f32, sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t
    sigs = {
        "kzd_version": (C.c_int, []),
        "kzd_last_error": (C.c_char_p, []),
        "kzd_device_info": (C.c_int, [C.POINTER(C.c_int)] * 3),
        "kzd_decode_infer": (C.c_int, [C.POINTER(vp), i32, C.POINTER(Level), i32, i32, i32, vp, vp]),
        "kzd_decode_dual": (C.c_int, [C.POINTER(vp), i32, C.POINTER(Level), i32, i32, i32, vp, vp, vp, vp]),
        "kzd_decode_filter": (C.c_int, [C.POINTER(vp), i32, C.POINTER(Level), i32, i32, i32, vp, vp, vp, f32, vp, vp, vp,
                                        i32, vp]),
        "kzd_decode_train": (C.c_int, [vp, i32, vp, vp, i32, i32, vp, vp, vp]),
        "kzd_decode_train_bwd": (C.c_int, [vp, i32, vp, i32, i32, vp, vp]),
        "kzd_assign_workspace_bytes": (sz, [i32, i32, i32]),
        "kzd_assign_report_offset": (sz, [i32, i32]),
        "kzd_tal_assign": (C.c_int, [vp] * 6 + [i32] * 5 + [f32] * 3 + [C.POINTER(Level), i32, i32, i32] + [vp] * 6 + [vp, sz, i32, vp]),
        "kzd_nms_filter": (C.c_int, [vp, i32, i32, i32, f32, i32, vp, i32, i32, vp, vp, vp, i32, vp]),
        "kzd_nms_append_labels": (C.c_int, [vp, vp, i32, i32, i32, i32, f32, i32, vp, i32, vp, vp, vp, i32, vp]),
        "kzd_nms_workspace_bytes": (sz, [i32, i32, i32]),
        "kzd_gt_count": (C.c_int, [vp, i32, i32, vp, vp, vp, vp]),
        "kzd_gt_pack": (C.c_int, [vp, i32, i32, i32, vp, vp, vp, vp, vp]),
        "kzd_loss_tail_fwd": (C.c_int, [vp] * 8 + [i32] * 3 + [vp, vp, vp]),
        "kzd_loss_tail_bwd": (C.c_int, [vp] * 8 + [i32] * 3 + [vp, vp, vp, vp]),
        "kzd_loss_tail_sparse_fwd": (C.c_int, [vp] * 9 + [i32] * 6 + [vp, vp, vp]),
        "kzd_loss_tail_sparse_bwd": (C.c_int, [vp] * 9 + [i32] * 6 + [vp, vp, vp, vp]),
        "kzd_scale_boxes": (C.c_int, [vp, vp, i32, i32, i32, vp, vp]),
        "kzd_match_workspace_bytes": (sz, [i32, i32, i32, i32]),
        "kzd_match_predictions": (C.c_int, [vp, vp, i32, i32, i32, vp, vp, i32, vp, i32, vp, vp, sz, vp]),
        "kzd_nms_suppress": (C.c_int, [vp, vp, vp, i32, i32, i32, f32, i32, i32, i32, f32, vp, vp, vp, sz, vp]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)  # AttributeError here = header and library out of sync
        fn.restype, fn.argtypes = res, args
    return lib, tuple(sigs)


lib, EXPORTS = _load()


def check(rc: int, what: str):
    if rc != 0:
        msg = lib.kzd_last_error().decode("utf-8", "replace")
        raise KzdError(f"{what} failed (code {rc}): {msg}")


def levels_array(levels):
    """``[(h, w, stride), ...]`` -> (ctypes array of kzd_level, total anchors)."""
    if len(levels) > KZD_MAX_LEVELS:
        raise KzdError(f"{len(levels)} pyramid levels > KZD_MAX_LEVELS={KZD_MAX_LEVELS}")
    arr = (Level * max(len(levels), 1))()
    off = 0
    for i, (h, w, s) in enumerate(levels):
        arr[i] = Level(int(h), int(w), off, float(s))
        off += int(h) * int(w)
    return arr, off


def dtype_code(dtype):
    """0 fp32, 1 fp16, 2 bf16 (the `in_dtype` / `score_dtype` arguments of include/kzd.h)."""
    import torch

    return {torch.float16: 1, torch.bfloat16: 2}.get(dtype, 0)


def require_cuda(t, name):
    if not t.is_cuda:
        raise KzdError(f"{name} must be a CUDA tensor: libkzd has no CPU path")
