#!/bin/bash
cd tools/exp/bin
for b in pl4c1 pl2b1; do
ncu --set full --clock-control none --import-source on -k regex:kern -s 3 -c 2 -o ../../../gpurun_out/prof_r02_plane_v7c_cheb_$b -f ./$b 0 1 > ../../../gpurun_out/ncu3_$b.log 2>&1
done
