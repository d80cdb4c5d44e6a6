#!/usr/bin/env bash
# TEST INFRASTRUCTURE.  Compiles the reference's OWN correlation extension (its only native code) from the sources where they
# lie under /root/reference into oracle/_ref/correlation_cuda.so (git-ignored, shipped to the GPU box by gpurun), for sm_100a.
# No reference source is copied or modified; ref_correlation_shim.h is force-included for PyTorch-header compatibility.
# Used by tests/test_correlation_gpu.py (checker) and bench.py's `correlation` side measurement (reference timed beside ours).
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF=${ACFM_REFERENCE_ROOT:-/root/reference}/multiframe/data/optical_flow/model/correlation_package
[[ -f $REF/correlation_cuda_kernel.cu ]] || { echo "reference sources not present: skipping oracle/_ref"; exit 0; }
OUT=$HERE/_ref
mkdir -p "$OUT"
if [[ -f $OUT/correlation_cuda.so && $OUT/correlation_cuda.so -nt $REF/correlation_cuda_kernel.cu && $OUT/correlation_cuda.so -nt $HERE/ref_correlation_shim.h ]]; then
  echo "oracle/_ref/correlation_cuda.so is up to date"; exit 0
fi
PY=${PYTHON:-python}
TINC=$($PY -c "import torch.utils.cpp_extension as c; print(' '.join('-I' + p for p in c.include_paths()))" 2>/dev/null)
TLIB=$($PY -c "import torch.utils.cpp_extension as c; print(c.library_paths()[0])" 2>/dev/null)
PYINC=$($PY -c "import sysconfig; print(sysconfig.get_paths()['include'])")
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
DEFS="-DTORCH_EXTENSION_NAME=correlation_cuda -DTORCH_API_INCLUDE_EXTENSION_H"
$NVCC -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -ccbin /usr/bin/g++ -Xcompiler -fPIC -w $DEFS $TINC -I"$PYINC" -I"$REF" \
      --expt-relaxed-constexpr -include "$HERE/ref_correlation_shim.h" -c "$REF/correlation_cuda_kernel.cu" -o "$OUT/ref_corr_kernel.o"
/usr/bin/g++ -O2 -std=c++17 -fPIC -w $DEFS $TINC -I"$PYINC" -I"$REF" -I/usr/local/cuda/include -include "$HERE/ref_correlation_shim.h" \
      -c "$REF/correlation_cuda.cc" -o "$OUT/ref_corr_host.o"
/usr/bin/g++ -shared -o "$OUT/correlation_cuda.so" "$OUT/ref_corr_host.o" "$OUT/ref_corr_kernel.o" -L"$TLIB" -lc10 -lc10_cuda -ltorch_cpu \
      -ltorch_cuda -ltorch -ltorch_python -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,"$TLIB"
rm -f "$OUT/ref_corr_host.o" "$OUT/ref_corr_kernel.o"
echo "built $OUT/correlation_cuda.so"
