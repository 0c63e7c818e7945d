"""TEST INFRASTRUCTURE ONLY - loader for the unmodified reference (vendored Ultralytics 8.3.63).

The reference is plain Python. In the build container it is imported where it lies
(``/root/reference/yolov12``); ``oracle/make_ref.py`` (run by ``__graft_entry__.build()``) stages a
byte-identical copy under the git-ignored ``oracle/_ref/yolov12`` that travels to the GPU box, so the
unmodified reference is importable there too (CPU and CUDA). ``available()`` is False only when
neither exists; callers then fall back to ``oracle.torch_oracle`` + ``tests/golden``.

Import recipe: three modules the reference imports at top level are
absent from this image (``matplotlib``, ``matplotlib.pyplot`` - ultralytics/utils/__init__.py:23,
utils/metrics.py:8; ``thop`` - utils/torch_utils.py:15) and are replaced by inert stubs.
"""
from __future__ import annotations

import importlib
import importlib.machinery
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOTS = (os.environ.get("KZD_REF", ""), "/root/reference/yolov12", os.path.join(_HERE, "_ref", "yolov12"))


def ref_root() -> str | None:
    for r in REF_ROOTS:
        if r and os.path.isdir(os.path.join(r, "ultralytics")):
            return r
    return None


def available() -> bool:
    return ref_root() is not None


class _Inert(types.ModuleType):
    """A module whose every attribute is a callable that returns itself."""

    def __init__(self, name):
        super().__init__(name)
        self.__path__ = []
        self.__spec__ = importlib.machinery.ModuleSpec(name, None, is_package=True)

    def __getattr__(self, item):
        if item.startswith("__"):
            raise AttributeError(item)
        return self

    def __call__(self, *a, **k):
        return self


_loaded = None


def load():
    """Return a namespace with the reference's hot-path symbols (imports once)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    root = ref_root()
    if root is None:
        raise RuntimeError("reference tree not present (expected /root/reference/yolov12 or oracle/_ref/yolov12)")
    os.environ.setdefault("YOLO_CONFIG_DIR", "/tmp/ultra_cfg")
    keep_omp = os.environ.get("OMP_NUM_THREADS")
    for name in ("matplotlib", "matplotlib.pyplot", "thop"):
        try:
            importlib.import_module(name)
        except Exception:
            sys.modules[name] = _Inert(name)
    if root not in sys.path:
        sys.path.insert(0, root)
    import torch

    import contextlib

    nthreads = torch.get_num_threads()
    with contextlib.redirect_stdout(sys.stderr):  # the package prints at import time; bench.py's stdout is one JSON line
        tal = importlib.import_module("ultralytics.utils.tal")
        ops = importlib.import_module("ultralytics.utils.ops")
        loss = importlib.import_module("ultralytics.utils.loss")
        metrics = importlib.import_module("ultralytics.utils.metrics")
        head = importlib.import_module("ultralytics.nn.modules.head")
        block = importlib.import_module("ultralytics.nn.modules.block")
    # ultralytics/__init__.py:8-9 forces OMP_NUM_THREADS=1 when unset; undo for our timing.
    if keep_omp is None:
        os.environ.pop("OMP_NUM_THREADS", None)
    torch.set_num_threads(nthreads)
    ns = types.SimpleNamespace(
        root=root, tal=tal, ops=ops, loss=loss, metrics=metrics, head=head, block=block,
        TaskAlignedAssigner=tal.TaskAlignedAssigner, make_anchors=tal.make_anchors,
        dist2bbox=tal.dist2bbox, bbox_iou=metrics.bbox_iou, Detect=head.Detect, DFL=block.DFL,
        non_max_suppression=ops.non_max_suppression, v8DetectionLoss=loss.v8DetectionLoss,
    )
    _loaded = ns
    return ns
