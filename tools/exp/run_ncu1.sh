#!/bin/bash
# ncu capture of the plane kernel (pl4ref, Q4 8x4 tile), APPLY mode launches only
cd tools/exp/bin
ncu --set full --clock-control none --import-source on -k regex:kern -c 3 -o ../../../gpurun_out/prof_r02_plane_v7a_q4 -f ./pl4ref 0 1 > ../../../gpurun_out/ncu1.log 2>&1
tail -5 ../../../gpurun_out/ncu1.log
