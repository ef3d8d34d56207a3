#!/bin/bash
# tools/sanitize.sh [memcheck|racecheck|synccheck|initcheck ...]   (default: memcheck racecheck)
# compute-sanitizer over the GPU parity tests on small meshes (run on a GPU box from the repo root, e.g.
#   gpurun --timeout 1500 -- 'tools/sanitize.sh > gpurun_out/sanitize.txt 2>&1').
# The selection keeps every apply kernel (plane-per-step, line-marching, cell-tile, variable coefficient, 2-D), the transfers,
# the fused smoother steps and one small V-cycle + CG solve; the BASELINE-size tests are left out (the tools slow kernels
# down 10-100 x).  Exit code 0 = every tool reported 0 errors.
set -u
cd "$(dirname "$0")/.."
TOOLS=${*:-memcheck racecheck}
SEL='test_vmult_matches_oracle or test_vmult_mixed_boundary or test_fused_smoother or test_transfers or test_diagonal_and_el or (test_vcycle_and_cg_match_oracle and h-2-8) or test_variable_coefficient_apply or test_2d'
rc=0
for tool in $TOOLS; do
  echo "=== compute-sanitizer --tool $tool"
  extra=""
  [ "$tool" = memcheck ] && extra="--leak-check full"
  timeout 1400 compute-sanitizer --tool "$tool" $extra --error-exitcode 77 --target-processes all \
    python -m pytest tests -m gpu -x -q -k "$SEL" -p no:cacheprovider 2>&1 | tail -25
  r=${PIPESTATUS[0]}
  echo "=== $tool exit code $r"
  [ "$r" -ne 0 ] && rc=1
done
exit $rc
