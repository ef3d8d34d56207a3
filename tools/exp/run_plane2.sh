#!/bin/bash
cd tools/exp/bin
for b in pl*; do [ -x ./$b ] && timeout 120 ./$b 0 10; done
echo "--- chunk sweeps"
for c in 4 8 16; do ./pl4a 0 10 $c | grep "mode="; done
for c in 4 7 10; do ./pl4c 0 10 $c | grep "mode=3"; done
for c in 4 8 16 29; do ./pl2a 0 10 $c | grep "mode="; done
for c in 4 8 16 ; do ./pl3a 0 10 $c | grep "mode="; done
