"""TEST INFRASTRUCTURE ONLY - recipe that stages the UNMODIFIED reference next to the oracle.

The reference's implementation of this path is pure Python (its vendored Ultralytics 8.3.63 under
``/root/reference/yolov12/ultralytics``), so there is nothing to compile: "building" it means copying the
package, byte for byte, into the git-ignored ``oracle/_ref/yolov12/`` so that it travels to the GPU box
with the snapshot (like the built ``libkzd.so`` does) while staying out of the repository's history.
``__graft_entry__.build()`` runs this whenever ``/root/reference`` is present (the build container); the
GPU box only ever sees the staged copy. ``oracle/ref_loader.py`` imports it from there.

With the staged copy on the box
  * the ``-m gpu`` tests compare libkzd with the reference's own ``TaskAlignedAssigner`` /
    ``Detect._inference`` / ``non_max_suppression`` executed on the same CUDA device,
  * ``install()`` is exercised against the real package,
  * ``bench.py``'s CPU arm times the reference itself (``cpu_baseline.kind = "reference"``).

    python oracle/make_ref.py            # stage (idempotent)
    python oracle/make_ref.py --check    # verify the staged copy is byte-identical to the source
"""
from __future__ import annotations

import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/yolov12/ultralytics"
DST_ROOT = os.path.join(HERE, "_ref", "yolov12")
DST = os.path.join(DST_ROOT, "ultralytics")
KEEP_EXT = (".py", ".yaml", ".yml", ".json", ".txt")  # code + the cfg tables the package reads at import time


def _files(root):
    for d, dirs, files in os.walk(root):
        dirs[:] = sorted(x for x in dirs if x != "__pycache__")
        for f in sorted(files):
            if f.endswith(KEEP_EXT):
                yield os.path.relpath(os.path.join(d, f), root)


def stage(verbose=True) -> bool:
    """Copy the reference package; returns False when the source tree is absent (GPU box)."""
    if not os.path.isdir(SRC):
        return False
    n = 0
    for rel in _files(SRC):
        s, t = os.path.join(SRC, rel), os.path.join(DST, rel)
        if os.path.exists(t) and filecmp.cmp(s, t, shallow=False):
            continue
        os.makedirs(os.path.dirname(t), exist_ok=True)
        shutil.copyfile(s, t)
        n += 1
    stale = set(_files(DST)) - set(_files(SRC))
    for rel in stale:
        os.remove(os.path.join(DST, rel))
    with open(os.path.join(DST_ROOT, "PROVENANCE"), "w") as fh:
        fh.write(f"unmodified copy of {SRC} made by oracle/make_ref.py; not product source, never committed\n")
    if verbose:
        print(f"oracle/_ref: {n} files staged, {len(stale)} stale files removed")
    return True


def check() -> bool:
    if not (os.path.isdir(SRC) and os.path.isdir(DST)):
        return False
    a, b = list(_files(SRC)), list(_files(DST))
    return a == b and all(filecmp.cmp(os.path.join(SRC, r), os.path.join(DST, r), shallow=False) for r in a)


if __name__ == "__main__":
    if "--check" in sys.argv:
        ok = check()
        print("oracle/_ref identical to the reference" if ok else "oracle/_ref differs from / lacks the reference")
        sys.exit(0 if ok else 1)
    sys.exit(0 if stage() else 1)
