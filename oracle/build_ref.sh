#!/usr/bin/env bash
# Builds the UNMODIFIED reference (sxyu/svox2) into oracle/_ref/ so that tests and
# bench.py can use it as (a) the pure-PyTorch CPU oracle and (b) the reference CUDA
# kernels compiled for sm_100 ("the kernel to beat") on the GPU box.
#
# Test infrastructure only.  Nothing under svox2_b200/ imports oracle/_ref.
# The reference sources are never copied into git history: oracle/_ref/ is git-ignored
# (it still travels to the GPU box with gpurun snapshots).
#
# Usage: oracle/build_ref.sh [/root/reference]
set -euo pipefail
REF=${1:-/root/reference}
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
if [ ! -d "$REF/svox2" ]; then
  echo "reference tree not found at $REF (expected on the build container only)"; exit 0
fi
# The reference's own callers (opt/opt.py, opt/render_imgs.py, opt/util/*): run UNMODIFIED by the
# drop-in tests against compat/svox2 and against the reference build (tests/test_gpu_dropin.py).
if [ ! -f "$OUT/opt/opt.py" ]; then
  mkdir -p "$OUT/opt"
  cp "$REF"/opt/*.py "$OUT/opt/"
  cp -r "$REF"/opt/util "$REF"/opt/configs "$OUT/opt/"
  find "$OUT/opt" -name __pycache__ -prune -exec rm -rf {} +
  chmod -R u+w "$OUT/opt"
  echo "reference callers installed into $OUT/opt"
fi
if ls "$OUT"/svox2/csrc*.so >/dev/null 2>&1; then
  echo "oracle/_ref already built"; exit 0
fi
WORK=$(mktemp -d /tmp/svox2_refbuild.XXXXXX)
cp -r "$REF"/. "$WORK"/
chmod -R u+w "$WORK"
( cd "$WORK" && TORCH_CUDA_ARCH_LIST=10.0 MAX_JOBS=${MAX_JOBS:-8} python setup.py build_ext --inplace > "$WORK/build.log" 2>&1 ) \
  || { echo "reference build failed, see $WORK/build.log"; exit 1; }
mkdir -p "$OUT/svox2"
cp "$WORK"/svox2/*.py "$OUT/svox2/"
cp "$WORK"/svox2/csrc*.so "$OUT/svox2/"
echo "reference installed into $OUT"
