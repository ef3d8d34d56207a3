#!/bin/bash
# GPU pass over plane-kernel tile variants (run on the GPU box from the repo root): every binary in tools/exp/bin named pl*
cd tools/exp/bin
for b in pl*; do
  [ -x ./$b ] && timeout 120 ./$b 0 10
done
