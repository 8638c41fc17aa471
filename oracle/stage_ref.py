"""TEST / MEASUREMENT INFRASTRUCTURE ONLY -- stage the UNMODIFIED reference for the GPU box.

    python oracle/stage_ref.py            # build container only (/root/reference is mounted here)

Copies the reference's own Python package (`/root/reference/waterEntropy`, byte for byte) into
`oracle/_ref/`, which is git-ignored (never part of the history) but travels to the GPU box with the
snapshot like a built `.so`.  `bench.py` imports it there - on top of `oracle/mda_shim`, the stand-in for
the uninstallable MDAnalysis - to time the reference itself (`cpu_baseline.reference_python`,
`kind: "reference"`) on its own arginine fixture, whose arrays are committed under `tests/golden/`.
Nothing under `waterentropy_b200/` imports it.
"""
from __future__ import annotations

import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/waterEntropy"
DST = os.path.join(HERE, "_ref", "waterEntropy")


def stage(verbose=True) -> bool:
    if not os.path.isdir(SRC):
        if verbose:
            print(f"{SRC} is not mounted: keeping whatever is staged under oracle/_ref")
        return os.path.isdir(DST)
    os.makedirs(os.path.dirname(DST), exist_ok=True)
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    shutil.copytree(SRC, DST, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    n = 0
    for root, _, files in os.walk(DST):
        for f in files:
            rel = os.path.relpath(os.path.join(root, f), DST)
            assert filecmp.cmp(os.path.join(root, f), os.path.join(SRC, rel), shallow=False), rel
            n += 1
    if verbose:
        print(f"staged {n} files of the unmodified reference into {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
