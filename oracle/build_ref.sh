#!/usr/bin/env bash
# TEST INFRASTRUCTURE — builds the *unmodified* reference C modules of the hot path
# (src_c/_extcoeff.c, src_c/vprofile.c) from where they lie under /root/reference into
# oracle/_ref/ (git-ignored; travels to the GPU box with the snapshot).
# Flags follow the reference's setup.py:19 (-O3 -ffast-math).  No source is copied.
set -euo pipefail
REF=${REF:-/root/reference}
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/_ref"
mkdir -p "$OUT"
if [ ! -d "$REF/src_c" ]; then
    echo "build_ref.sh: $REF/src_c not present; keeping prebuilt files in $OUT" >&2
    exit 0
fi
PY=${PYTHON:-python}
PYINC=$($PY -c "import sysconfig;print(sysconfig.get_paths()['include'])")
[ -f "$PYINC/Python.h" ] || PYINC=/usr/include/python3.12
NPINC=$($PY -c "import numpy;print(numpy.get_include())")
EXT=$($PY -c "import sysconfig;print(sysconfig.get_config_var('EXT_SUFFIX'))")
for m in _extcoeff vprofile; do
    gcc -shared -fPIC -O3 -ffast-math -w -I"$PYINC" -I"$NPINC" -I"$REF/src_c/include" \
        "$REF/src_c/$m.c" -o "$OUT/$m$EXT" -lm
done
# A scratch copy of the reference's PYTHON package with its own C modules, except the two the
# engine replaces (tests/test_gpu_dropin.py runs the unmodified reference on the GPU engine
# through pyratbay_b200/shim).  Git-ignored like the rest of oracle/_ref.
PKG="$OUT/shimmed/pyratbay"
rm -rf "$OUT/shimmed"
mkdir -p "$OUT/shimmed"
cp -r "$REF/pyratbay" "$PKG"
mkdir -p "$PKG/lib"
for f in "$REF"/src_c/*.c; do
    m=$(basename "$f" .c)
    case "$m" in _extcoeff|vprofile) continue;; esac
    gcc -shared -fPIC -O3 -ffast-math -w -I"$PYINC" -I"$NPINC" -I"$REF/src_c/include" \
        "$f" -o "$PKG/lib/$m$EXT" -lm
done
PYTHONPATH="$HERE/.." $PY -c "from pyratbay_b200.shim import install_into; install_into('$PKG/lib')"
echo "built: $(ls "$OUT")"
