#!/usr/bin/env bash
# Builds kliff_b200/csrc/libkliff_b200.so for sm_100a (cross-compiles without a GPU).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
OUT="${KB_OUT:-$HERE/libkliff_b200.so}"
SRCS=(scan.cu pad.cu neighbor.cu batch.cu symfun_general.cu symfun_tiled.cu symfun_warp.cu symfun_split.cu symfun_api.cu scatter.cu scatter_seg.cu moments.cu train.cu train_fast.cu peer.cu misc.cu)
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -shared
       -Xcompiler -fPIC,-ffp-contract=off,-fvisibility=default
       -I"$ROOT/include" -I"$HERE" ${KB_NVCC_EXTRA:-})
cd "$HERE"
"$NVCC" "${FLAGS[@]}" "${SRCS[@]}" -o "$OUT" -lcudart
echo "built $OUT"
