#!/usr/bin/env python
"""Copy the reference's Stage-1 Python sources into baseline/_ref/ (git-ignored, NOT gpurun-ignored).

    python baseline/install_reference.py            # no-op when /root/reference is absent (GPU box: uses the shipped copy)

Only what the Stage-1 path imports is copied (model/, CLIP/, loss/, utils/, dataset/, args.py, logger.py,
train_stage1.py, validate.py, demo.py); IRNet/, figs/, Stage-2 training and git metadata stay behind.
"""
from __future__ import annotations

import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("TRIS_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(HERE, "_ref")
KEEP_DIRS = ("model", "CLIP", "loss", "utils", "dataset")
KEEP_FILES = ("args.py", "logger.py", "train_stage1.py", "validate.py", "demo.py", "__init__.py", "LICENSE")


def install(verbose: bool = True) -> str | None:
    if not os.path.isdir(os.path.join(SRC, "model")):
        if verbose:
            print(f"[install_reference] {SRC} not present; keeping {DST if os.path.isdir(DST) else 'nothing'}")
        return DST if os.path.isdir(DST) else None
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(DST)
    for d in KEEP_DIRS:
        shutil.copytree(os.path.join(SRC, d), os.path.join(DST, d), ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    for f in KEEP_FILES:
        if os.path.exists(os.path.join(SRC, f)):
            shutil.copy2(os.path.join(SRC, f), os.path.join(DST, f))
    if verbose:
        n = sum(len(fs) for _, _, fs in os.walk(DST))
        print(f"[install_reference] copied {n} files from {SRC} to {DST}")
    return DST


if __name__ == "__main__":
    sys.exit(0 if install() else 1)
