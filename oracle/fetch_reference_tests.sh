#!/bin/bash
# Copies the reference's own unit tests (dakofler/Compyute tests/) next to the oracle as a git-ignored, read-only fixture so
# that tests/test_gpu_dropin.py can run them UNMODIFIED against the `compyute` import shim on the GPU box (where
# /root/reference does not exist).  Nothing under oracle/_ref/ is ever committed (.gitignore) or imported by the product.
set -e
SRC=${1:-/root/reference/tests}
DST=$(dirname "$0")/_ref/reference_tests
[ -d "$SRC" ] || { echo "no reference tests at $SRC"; exit 0; }
rm -rf "$DST"; mkdir -p "$DST"
cp -r "$SRC" "$DST/tests"
find "$DST" -name "__pycache__" -prune -exec rm -rf {} +
echo "copied $(find "$DST/tests" -name '*.py' | wc -l) files to $DST/tests"
