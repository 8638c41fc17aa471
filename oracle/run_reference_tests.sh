#!/bin/bash
# TEST INFRASTRUCTURE ONLY.  Runs the UNMODIFIED reference test-suite
# (/root/reference/tests, minus the SLURM/dask-jobqueue string tests) on top of
# oracle/mda_shim.  Only works in the build container, where /root/reference is
# mounted; it is the gate that pins the shim against the reference's goldens.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REPO="$(dirname "$HERE")"
export PYTHONPATH="$HERE/mda_shim:$REPO:/root/reference:$PYTHONPATH"
cd /root/reference
exec python -m pytest -p no:cacheprovider -q \
    tests/maths tests/analysis tests/utils/test_selections.py tests/recipes "$@"
