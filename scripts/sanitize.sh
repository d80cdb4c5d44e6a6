#!/usr/bin/env bash
# compute-sanitizer over the rasterizer tests (run on a B200 box through gpurun):
#   gpurun --timeout 1500 -- 'bash scripts/sanitize.sh r05'
# memcheck (out-of-bounds / misaligned), racecheck (shared-memory hazards: the warp slabs of raster_fwd_kernel alias the
# staging scratch and rely on named-barrier / __syncwarp ordering), initcheck (reads of uninitialised global memory) and
# synccheck.  Small shapes only: the tools slow kernels down 10-100x.  Summaries -> gpurun_out/sanitize_<tool>_<tag>.log;
# scripts/summarize_profiles.py does not touch them: copy the tail lines into profiles/sanitize_<tag>.md.
set -uo pipefail
tag=${1:-r05}
sel='golden_small or edge_cases or split_path or bounds_only or epsilon_setting or misaligned or fused_mask_losses or fused_visible'
mkdir -p gpurun_out
rc=0
for tool in memcheck racecheck initcheck synccheck; do
  compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 \
    python -m pytest tests/test_raster_gpu.py tests/test_deform_gpu.py -x -q -m gpu -k "$sel or handle" > gpurun_out/sanitize_${tool}_${tag}.log 2>&1
  r=$?
  echo "== $tool: exit $r"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/sanitize_${tool}_${tag}.log | tail -3
  [ $r -ne 0 ] && rc=$r
done
exit $rc
