#!/usr/bin/env bash
# Builds libacfm_b200.so in-tree for sm_100a (cross-compiles without a GPU).  Sources compile in parallel.
set -uo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC
       -ccbin /usr/bin/g++ --fmad=true -Xptxas -v)
SRCS=(api.cu project.cu skin.cu handle_solve.cu camera.cu raster_fwd.cu raster_bwd.cu shade.cu uvsample.cu losses.cu reproj.cu laplacian.cu targets.cu priors.cu correlation.cu)
OBJS=()
PIDS=()
NAMES=()
mkdir -p _obj
for s in "${SRCS[@]}"; do
  o=_obj/${s%.cu}.o
  if [[ ! -f $o || $s -nt $o || common.cuh -nt $o || raster_common.cuh -nt $o || ../../include/acfm_b200.h -nt $o ]]; then
    "$NVCC" "${FLAGS[@]}" -c "$s" -o "$o" 2> "_obj/${s%.cu}.ptxas.log" &
    PIDS+=($!)
    NAMES+=("$s")
  fi
  OBJS+=("$o")
done
fail=0
for i in "${!PIDS[@]}"; do
  if ! wait "${PIDS[$i]}"; then
    echo "nvcc failed on ${NAMES[$i]}:" >&2
    cat "_obj/${NAMES[$i]%.cu}.ptxas.log" >&2
    rm -f "_obj/${NAMES[$i]%.cu}.o"
    fail=1
  fi
done
[[ $fail == 0 ]] || exit 1
"$NVCC" -gencode arch=compute_100a,code=sm_100a -shared -o ../libacfm_b200.so "${OBJS[@]}" -ccbin /usr/bin/g++ -lcudart || exit 1
echo "built $(cd .. && pwd)/libacfm_b200.so"
