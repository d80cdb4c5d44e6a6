#!/usr/bin/env bash
# Profiling recipe of this repo (run on a B200 box through gpurun; one ncu mode per call):
#   gpurun -- 'bash scripts/profile_round.sh launches r01'     every launch of a short bench run with its device time
#   gpurun -- 'bash scripts/profile_round.sh full_fwd r01'     ncu --set full of the forward rasterizer launch of the bench step
#   gpurun -- 'bash scripts/profile_round.sh full_fill r01'    same for the concurrent padding kernel of the split forward path
#   gpurun -- 'bash scripts/profile_round.sh full_bwd r01'     same for the backward rasterizer
# Outputs land in gpurun_out/; scripts/summarize_profiles.py turns them into the committed profiles/*.md / traffic.json.
set -uo pipefail
mode=$1; tag=${2:-r01}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
mkdir -p gpurun_out
case $mode in
  launches)
    $CMD > gpurun_out/plain_${tag}.log 2>&1 &&
    ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_${tag}.csv $CMD > gpurun_out/ncu_launches_${tag}.log 2>&1
    ;;
  full_fwd)
    # launch 1 = target set-up render (64 renders), 2-4 = warm-up steps, 5 = first timed step (512 renders)
    $CMD > gpurun_out/plain_${tag}.log 2>&1 &&
    ncu --set full --clock-control none --import-source on -k regex:raster_fwd -s 4 -c 1 -o gpurun_out/prof_fwd_${tag} -f $CMD > gpurun_out/ncu_fwd_${tag}.log 2>&1
    ;;
  full_fill)
    # the concurrent padding kernel of the split forward path (serialised under ncu)
    $CMD > gpurun_out/plain_${tag}.log 2>&1 &&
    ncu --set full --clock-control none --import-source on -k regex:raster_fill -s 4 -c 1 -o gpurun_out/prof_fill_${tag} -f $CMD > gpurun_out/ncu_fill_${tag}.log 2>&1
    ;;
  full_bwd)
    $CMD > gpurun_out/plain_${tag}.log 2>&1 &&
    ncu --set full --clock-control none --import-source on -k regex:raster_soft_bwd -s 3 -c 1 -o gpurun_out/prof_bwd_${tag} -f $CMD > gpurun_out/ncu_bwd_${tag}.log 2>&1
    ;;
esac
rc=$?; echo "profile_round: $mode $tag exit $rc"; exit $rc
