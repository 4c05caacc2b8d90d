#!/usr/bin/env bash
# Build libreid_b200.so for sm_100a, in-tree (the .so travels to the GPU box with the snapshot).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="${HERE}/../libreid_b200.so"
BUILD="${HERE}/build"
EXTRA=()
if [[ "${REID_DEV:-0}" != "0" ]]; then        # developer build: -DREID_DEV enables the timing switches (common.cuh)
  OUT="${HERE}/../libreid_b200_dev.so"
  BUILD="${HERE}/build_dev"
  EXTRA=(-DREID_DEV)
fi
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC
       --expt-relaxed-constexpr -Xptxas -v)
mkdir -p "${BUILD}"
pids=()
objs=()
for src in "${HERE}"/*.cu; do
  obj="${BUILD}/$(basename "${src%.cu}").o"
  objs+=("${obj}")
  if [[ ! -f "${obj}" || "${src}" -nt "${obj}" || "${HERE}/common.cuh" -nt "${obj}" || "${HERE}/tc_ptx.cuh" -nt "${obj}" || "${HERE}/../../include/reid_b200.h" -nt "${obj}" ]]; then
    ( "${NVCC}" "${FLAGS[@]}" "${EXTRA[@]}" -c "${src}" -o "${obj}" > "${obj}.log" 2>&1 || { cat "${obj}.log"; exit 1; } ) &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [[ -n "${p}" ]] && wait "${p}"; done
"${NVCC}" -gencode arch=compute_100a,code=sm_100a -shared -o "${OUT}" "${objs[@]}" -lcuda
echo "built ${OUT}"
