#!/bin/bash
# Compiles the REFERENCE'S OWN GPU runner mains — tests/mnist/main-gpu.cpp, tests/fashion-mnist/main-gpu.cpp and
# tests/fashion-mnist/main_gpu_deep.cpp, byte for byte as they lie under $REF — against this backend's drop-in headers
# (include/compat mirrors the reference's src/ and tests/mnist/ include paths) and links them with libb200lbfgs.so.
# The mains include their headers by relative path, so they are staged (in a temporary directory, never in the repo) at the
# position tests/<dir>/ relative to a copy of include/compat. Output: examples/_ref/<name> (git-ignored; travels to the GPU box).
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(dirname "$HERE")"
REF="${REF:-/root/reference}"
OUT="$HERE/_ref"
[ -d "$REF/tests" ] || { echo "no reference checkout at $REF: nothing to do"; exit 0; }
TMP="$(mktemp -d)"
trap 'rm -rf "$TMP"' EXIT
cp -r "$ROOT/include" "$TMP/include"
mkdir -p "$OUT" "$TMP/include/compat/tests/fashion-mnist"
for src in mnist/main-gpu.cpp fashion-mnist/main-gpu.cpp fashion-mnist/main_gpu_deep.cpp; do
  name="$(echo "$src" | sed 's#/#_#; s#\.cpp$##; s#-#_#g')"
  cp "$REF/tests/$src" "$TMP/include/compat/tests/$src"
  /usr/bin/g++ -O2 -std=c++17 -fopenmp -o "$OUT/$name" "$TMP/include/compat/tests/$src" \
      -L"$ROOT/lbfgs_ffnn_b200" -lb200lbfgs -Wl,-rpath,'$ORIGIN/../../lbfgs_ffnn_b200'
  echo "built examples/_ref/$name from $REF/tests/$src (unmodified)"
done
