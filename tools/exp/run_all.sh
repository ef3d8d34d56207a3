#!/bin/bash
# run every experiment binary: tools/exp/run_all.sh CELLS [REPS]
cd "$(dirname "$0")/bin" || exit 1
for b in $(ls | grep -v '\.o$'); do timeout 60 ./$b ${1:-116} ${2:-10} || echo "$b failed"; done
