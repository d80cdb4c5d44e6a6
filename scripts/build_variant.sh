#!/usr/bin/env bash
# Builds the current sources into .variants/lib_<name>.so WITH the tuning hooks (-DACFM_TUNING: the ACFM_* environment
# variables of raster_fwd.cu / correlation.cu); the shipped libacfm_b200.so (csrc/build.sh) is built without them.
# usage: scripts/build_variant.sh <name> [extra nvcc flags...]        then: LIB=.variants/lib_<name>.so python scripts/time_fwd.py ...
set -euo pipefail
cd "$(dirname "$0")/.."
name=$1; shift
src=acfm_video_3d_reconstruction_b200/csrc
out=.variants/obj_$name
mkdir -p "$out"
pids=()
for s in $src/*.cu; do
  b=$(basename "${s%.cu}")
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -ccbin /usr/bin/g++ \
    --fmad=true -DACFM_TUNING "$@" -c "$s" -o "$out/$b.o" 2> "$out/$b.log" &
  pids+=($!)
done
for p in "${pids[@]}"; do wait "$p"; done
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o .variants/lib_$name.so "$out"/*.o -ccbin /usr/bin/g++ -lcudart
echo "built .variants/lib_$name.so"
