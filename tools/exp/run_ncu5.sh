#!/bin/bash
cd tools/exp/bin
ncu --section SourceCounters --section SpeedOfLight --section WarpStateStats --clock-control none --import-source on -k regex:^kern -c 1 -o ../../../gpurun_out/prof_r02_plane_v7e_pl4u -f ./pl4u 64 1 8 > ../../../gpurun_out/ncu5.log 2>&1
tail -3 ../../../gpurun_out/ncu5.log
