#!/usr/bin/env bash
# Builds baseline/_ref: the UNMODIFIED reference package (copied from $REF, its pybind11 extensions
# compiled in place by its own setup.py) for bench.py's reference arm.  SURVEY.md Appendix C.
# baseline/_ref is git-ignored; it travels to the GPU box with the gpurun snapshot.
# Needs: $REF (default /root/reference), pybind11.  No network.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${REF:-/root/reference}"
if [ ! -d "$REF/kliff" ]; then
  echo "reference tree $REF not present: keeping prebuilt baseline/_ref (if any)"; exit 0
fi
SO=$(ls "$HERE"/_ref/kliff_src/kliff/legacy/descriptors/symmetry_function/sf*.so 2>/dev/null | head -1 || true)
if [ -n "$SO" ] && [ "${1:-}" != "--force" ]; then echo "baseline/_ref already built"; exit 0; fi
rm -rf "$HERE/_ref/kliff_src"
mkdir -p "$HERE/_ref"
cp -r "$REF" "$HERE/_ref/kliff_src"
chmod -R u+w "$HERE/_ref/kliff_src"
(cd "$HERE/_ref/kliff_src" && python setup.py build_ext --inplace -j 8 > "$HERE/_ref/build.log" 2>&1)
rm -rf "$HERE/_ref/kliff_src/build" "$HERE/_ref/kliff_src/.git" "$HERE/_ref/kliff_src/docs"
echo "built baseline/_ref (reference $(cd "$REF" && git rev-parse --short HEAD 2>/dev/null || echo '?'))"
