#!/bin/bash
cd tools/exp/bin
for b in pl2a pl4c pl3a; do timeout 120 ./$b 0 10 | grep "mode=\|rel"; done
for b in pl2a pl4c; do
ncu --set full --clock-control none --import-source on -k regex:^kern -c 4 -o ../../../gpurun_out/prof_r02_plane_v7d_$b -f ./$b 0 1 > ../../../gpurun_out/ncu4_$b.log 2>&1
done
