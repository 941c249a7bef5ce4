#!/usr/bin/env python
"""Stages the UNMODIFIED reference sources under baseline/_ref/ (git-ignored, NOT gpurun-ignored) so that they travel to
the GPU box, where /root/reference does not exist.  Nothing under baseline/_ref is repo code: it is the reference's own
`src/` tree, byte for byte, used by

  * tests/test_gpu_reference_callers.py - the reference's own train.py / trainCas*.py driving this package's kernels;
  * bench.py --impl reference          - the reference's own CPU implementation (kind "reference");
  * bench.py --impl torch-gpu          - the incumbent: the same unmodified modules through stock PyTorch/cuDNN on the B200.

The reference is plain Python without a setup.py / pyproject.toml, so `pip install --target baseline/_ref /root/reference`
cannot work (recorded in DESIGN.md); copying the tree is the equivalent installation."""
import filecmp
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("SRCGAN_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")


def stage() -> bool:
    if not os.path.isfile(os.path.join(SRC, "src", "train.py")):
        return False
    os.makedirs(DST, exist_ok=True)
    for sub in ("src",):
        d = os.path.join(DST, sub)
        if os.path.isdir(d):
            shutil.rmtree(d)
        shutil.copytree(os.path.join(SRC, sub), d, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    for f in ("LICENSE", "README.md"):
        if os.path.isfile(os.path.join(SRC, f)):
            shutil.copy2(os.path.join(SRC, f), os.path.join(DST, f))
    cmp = filecmp.dircmp(os.path.join(SRC, "src"), os.path.join(DST, "src"), ignore=["__pycache__"])
    assert not cmp.diff_files and not cmp.left_only, (cmp.diff_files, cmp.left_only)
    return True


if __name__ == "__main__":
    ok = stage()
    print("staged %s/src -> %s" % (SRC, DST) if ok else "reference tree not present at %s: nothing staged" % SRC)
    sys.exit(0)
