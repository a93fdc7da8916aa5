#!/usr/bin/env bash
# Build the reference's ONLY native component -- tt_sketch/drm/fast_lazy_gaussian.pyx --
# from the sources where they lie under /root/reference into oracle/_ref/ (git-ignored,
# travels to the GPU box).  Nothing is copied into the repo: cython reads the .pyx in
# place and writes the generated C + the extension module only under oracle/_ref/.
#
# TEST INFRASTRUCTURE ONLY: the result is used by tests/ and bench.py's cpu_baseline /
# --impl reference leg to pin oracle/lazy_gaussian.c and oracle/sketch_oracle.py.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${TTSK_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
PYX="$REF/tt_sketch/drm/fast_lazy_gaussian.pyx"
if [ ! -f "$PYX" ]; then
    echo "build_ref: $PYX not present (GPU box?) -- keeping prebuilt files in $OUT" >&2
    exit 0
fi
mkdir -p "$OUT"
PY="${PYTHON:-python}"
NPINC="$($PY -c 'import numpy; print(numpy.get_include())')"
PYINC="$($PY -c 'import sysconfig; print(sysconfig.get_paths()["include"])')"
SUFFIX="$($PY -c 'import sysconfig; print(sysconfig.get_config_var("EXT_SUFFIX"))')"
$PY -m cython -3 "$PYX" -o "$OUT/fast_lazy_gaussian.c"
# same flags as the reference's setup.py:14-15 (-fopenmp; no prange is used so it is serial)
gcc -O2 -fPIC -shared -fopenmp -fwrapv -fno-strict-aliasing \
    -DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION \
    -I"$NPINC" -I"$PYINC" "$OUT/fast_lazy_gaussian.c" -o "$OUT/fast_lazy_gaussian$SUFFIX"
rm -f "$OUT/fast_lazy_gaussian.c"
echo "build_ref: built $OUT/fast_lazy_gaussian$SUFFIX"
# The reference itself is pure Python around that one extension: install the package (an unmodified copy of
# its tt_sketch/ directory, like `pip install --target`) next to it so that bench.py --impl reference and the
# cpu_baseline leg can time THE REFERENCE on the GPU box's host cores.  oracle/_ref/ is git-ignored build
# output (it travels with gpurun); nothing of it enters the repository.
PKG="$OUT/pkg"
rm -rf "$PKG"
mkdir -p "$PKG"
cp -r "$REF/tt_sketch" "$PKG/tt_sketch"
find "$PKG" -name "__pycache__" -type d -exec rm -rf {} + 2>/dev/null || true
cp "$OUT/fast_lazy_gaussian$SUFFIX" "$PKG/tt_sketch/drm/"
echo "build_ref: installed the reference package under $PKG"
