#!/bin/bash
# tools/exp/build_exp.sh NAME P BX BY LZ NT MINB [extra -D flags]   -> tools/exp/bin/NAME
# EXP_SRC=exp_pipe.cu builds the pipelined variant (NT = threads per group, the CTA has twice as many)
set -e
cd /root/repo/tools/exp; mkdir -p bin
NAME=$1; P=$2; BX=$3; BY=$4; LZ=$5; NT=$6; MINB=$7; shift 7
[ -f bin/pmg_fe.o ] || gcc -O2 -c -I../../include -I../../portable-multigrid_b200/host -I/usr/local/cuda/include ../../portable-multigrid_b200/host/pmg_fe.c -o bin/pmg_fe.o
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -I../../portable-multigrid_b200/csrc \
  -DC_NAME=$NAME -DC_P=$P -DC_BX=$BX -DC_BY=$BY -DC_LZ=$LZ -DC_NT=$NT -DMINB=$MINB "$@" ${EXP_SRC:-exp_sweep.cu} bin/pmg_fe.o -o bin/$NAME
