#!/bin/bash
# tools/exp/build_plane.sh NAME P BX BY NT MINB [extra -D flags]   -> tools/exp/bin/NAME   (harness: exp_plane.cu)
set -e
cd /root/repo/tools/exp; mkdir -p bin
NAME=$1; P=$2; BX=$3; BY=$4; NT=$5; MINB=$6; shift 6
[ -f bin/pmg_fe.o ] || gcc -O2 -c -I../../include -I../../portable-multigrid_b200/host -I/usr/local/cuda/include ../../portable-multigrid_b200/host/pmg_fe.c -o bin/pmg_fe.o
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -I../../portable-multigrid_b200/csrc \
  -DC_NAME=$NAME -DC_P=$P -DC_BX=$BX -DC_BY=$BY -DC_NT=$NT -DMINB=$MINB "$@" exp_plane.cu bin/pmg_fe.o -o bin/$NAME
