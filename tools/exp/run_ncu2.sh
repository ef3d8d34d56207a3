#!/bin/bash
cd tools/exp/bin
for b in pl4c pl2ref; do
ncu --set full --clock-control none --import-source on -k regex:kern -c 2 -o ../../../gpurun_out/prof_r02_plane_v7b_$b -f ./$b 0 1 > ../../../gpurun_out/ncu2_$b.log 2>&1
done
