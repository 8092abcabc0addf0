#!/bin/bash
# TEST INFRASTRUCTURE.  Builds oracle/_ref/libemd_ref.so: the REFERENCE's own EMD auction kernels
# (/root/reference/modules/loss/emd/emd_cuda.cu:9-226 and :284-300, device code only) behind the raw-pointer harness
# oracle/emd_ref_harness.cu.  The kernel text is extracted into a temporary directory for the compile and removed
# again: nothing of the reference is written into the repository, only the binary lands in oracle/_ref/ (git-ignored,
# shipped to the GPU box by gpurun).  Skips quietly when the reference tree is absent (the GPU box uses the prebuilt .so).
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${VPN_REFERENCE_ROOT:-/root/reference}/modules/loss/emd/emd_cuda.cu"
OUT="$HERE/_ref/libemd_ref.so"
if [ ! -f "$REF" ]; then
  echo "build_ref_emd: $REF not present - keeping any prebuilt $OUT"; exit 0
fi
if [ -f "$OUT" ] && [ "$OUT" -nt "$REF" ] && [ "$OUT" -nt "$HERE/emd_ref_harness.cu" ] && [ "$OUT" -nt "$0" ]; then
  exit 0
fi
TMP="$(mktemp -d)"; trap 'rm -rf "$TMP"' EXIT
sed -n '9,226p;284,300p' "$REF" > "$TMP/emd_ref_kernels.cuh"
mkdir -p "$HERE/_ref"
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -shared -Xcompiler -fPIC \
     -DEMD_REF_KERNELS="\"$TMP/emd_ref_kernels.cuh\"" "$HERE/emd_ref_harness.cu" -o "$OUT" -lcudart
echo "build_ref_emd: built $OUT"
