#!/bin/bash
# compute-sanitizer over the small-mesh GPU paths (run on the GPU box):
#   tools/sanitize.sh [outdir]       -> <outdir>/sanitize_<tool>_<case>.log + summary.txt
# Tools: memcheck, racecheck (shared-memory hazards: the stiffness kernel relies on split mbarrier
# arrive/wait, cp.async and per-warp tile exchange), synccheck (barrier misuse), initcheck.
# Cases: every stiffness kernel variant (regular bricks, generic via WFX_REGULAR=0, affine, mixed,
# interface/interior parts, cell-colour kernel), P2/P4/P6, fp64/fp32, the RK4 step with and without
# graph replay, geometry / mass / boundary set-up kernels.
out=${1:-gpurun_out/sanitize}
mkdir -p "$out"
summary="$out/summary.txt"
: > "$summary"
run() { # tool case env...
  tool=$1; case=$2; shift 2
  log="$out/sanitize_${tool}_${case}.log"
  env "$@" timeout 900 compute-sanitizer --tool "$tool" --print-limit 20 --error-exitcode 9 \
      python tools/sanitize_cases.py "$case" > "$log" 2>&1
  rc=$?
  errs=$(grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard" "$log" | tail -3 | tr '\n' ' ')
  echo "$tool $case rc=$rc $errs" | tee -a "$summary"
}
for case in regular generic affine mixed parts cell p2 p6 fp32 rk4 rk4_nograph setup; do
  extra=""
  [ "$case" = generic ] && extra="WFX_REGULAR=0"
  [ "$case" = rk4_nograph ] && extra="WFX_WAVE_GRAPH=0"
  run memcheck "$case" WFX_SAN=1 $extra
done
for case in regular generic affine mixed p2 p6 fp32; do
  extra=""
  [ "$case" = generic ] && extra="WFX_REGULAR=0"
  run racecheck "$case" WFX_SAN=1 $extra
  run synccheck "$case" WFX_SAN=1 $extra
done
run initcheck regular WFX_SAN=1
echo "done" >> "$summary"
