#!/bin/bash
# Pipelined apply kernel (csrc/pmg_apply_sweep_pipe.h): build, time and self-check against the plain kernel -> gpurun_out/exp_pipe.txt
#   gpurun --timeout 300 -- 'bash tools/exp/run_pipe.sh'
cd "$(dirname "$0")" || exit 1
mkdir -p bin ../../gpurun_out
export EXP_SRC=exp_pipe.cu
[ -x bin/pipe_q4_f0 ] || {
  ./build_exp.sh pipe_q4_f0 4 4 4 1 128 2 -DC_US=1 -DC_FM=0 -DC_RL=1
  ./build_exp.sh pipe_q4_f3 4 4 4 1 128 1 -DC_US=1 -DC_FM=3 -DC_RL=1
  ./build_exp.sh pipe_q2_f0 2 8 8 2 128 2 -DC_US=1 -DC_FM=0
  ./build_exp.sh pipe_q2_f3 2 8 8 2 128 2 -DC_US=1 -DC_FM=3
  ./build_exp.sh pipe_q3_f0 3 6 5 1 128 2 -DC_US=0 -DC_FM=0
  ./build_exp.sh pipe_q3_f3 3 6 5 1 128 2 -DC_US=0 -DC_FM=3
  ./build_exp.sh pipe_q1_f0 1 16 16 4 128 2 -DC_US=0 -DC_FM=0
  # fused step with b / x_old read from global memory by the z sweep (EG = 1): 2 CTAs/SM for the pipelined kernel ...
  ./build_exp.sh pipe_eg_q4_f3 4 4 4 1 128 2 -DC_US=1 -DC_FM=3 -DC_RL=1 -DC_EG=1
  ./build_exp.sh pipe_eg_q2_f3 2 8 8 2 128 2 -DC_US=1 -DC_FM=3 -DC_EG=1
  ./build_exp.sh pipe_eg_q3_f3 3 6 5 1 128 2 -DC_US=0 -DC_FM=3 -DC_EG=1
  # ... and 4 CTAs/SM for the plain kernel (exp_sweep.cu: timing only), next to the shipped configuration
  EXP_SRC=exp_sweep.cu ./build_exp.sh eg_q4_f3_m4 4 4 4 1 128 4 -DC_US=1 -DC_FM=3 -DC_RL=1 -DC_EG=1
  EXP_SRC=exp_sweep.cu ./build_exp.sh eg_q2_f3_m4 2 8 8 2 128 4 -DC_US=1 -DC_FM=3 -DC_RL=1 -DC_EG=1
  EXP_SRC=exp_sweep.cu ./build_exp.sh eg_q4_f3_m3 4 4 4 1 128 3 -DC_US=1 -DC_FM=3 -DC_EG=1
  EXP_SRC=exp_sweep.cu ./build_exp.sh o_q4_base_f3 4 4 4 1 128 3 -DC_US=1 -DC_FM=3
  EXP_SRC=exp_sweep.cu ./build_exp.sh a_q4_base_f0 4 4 4 1 128 3 -DC_US=1 -DC_FM=0
  # the shipped Q4 apply (rolled loops, 4 CTAs/SM) without / with one mbarrier arrival per warp instead of per thread
  EXP_SRC=exp_sweep.cu ./build_exp.sh c_q4_rl_m4_f0 4 4 4 1 128 4 -DC_US=1 -DC_FM=0 -DC_RL=1
  EXP_SRC=exp_sweep.cu ./build_exp.sh wa_q4_rl_m4_f0 4 4 4 1 128 4 -DC_US=1 -DC_FM=0 -DC_RL=1 -DPMG_SWEEP_WARP_ARRIVE
  EXP_SRC=exp_sweep.cu ./build_exp.sh wa_q4_f3 4 4 4 1 128 3 -DC_US=1 -DC_FM=3 -DPMG_SWEEP_WARP_ARRIVE
  # ... and with the loader warp's row loop relieved of its per-copy address arithmetic (12 -> 3.5 instructions per cp.async)
  EXP_SRC=exp_sweep.cu ./build_exp.sh lean_q4_rl_m4_f0 4 4 4 1 128 4 -DC_US=1 -DC_FM=0 -DC_RL=1 -DPMG_SWEEP_LEAN_LOADER
  EXP_SRC=exp_sweep.cu ./build_exp.sh lean_wa_q4_rl_m4_f0 4 4 4 1 128 4 -DC_US=1 -DC_FM=0 -DC_RL=1 -DPMG_SWEEP_LEAN_LOADER -DPMG_SWEEP_WARP_ARRIVE
  EXP_SRC=exp_sweep.cu ./build_exp.sh lean_q2_f0 2 8 8 2 128 3 -DC_US=1 -DC_FM=0 -DPMG_SWEEP_LEAN_LOADER
  ./build_exp.sh pipe_wa_q4_f0 4 4 4 1 128 2 -DC_US=1 -DC_FM=0 -DC_RL=1 -DPMG_SWEEP_WARP_ARRIVE
}
O=../../gpurun_out/exp_pipe.txt
for b in a_q4_base_f0 c_q4_rl_m4_f0 wa_q4_rl_m4_f0 lean_q4_rl_m4_f0 lean_wa_q4_rl_m4_f0 lean_q2_f0 pipe_q4_f0 pipe_wa_q4_f0 wa_q4_f3 o_q4_base_f3 pipe_q4_f3 pipe_eg_q4_f3 eg_q4_f3_m4 eg_q4_f3_m3 pipe_q2_f0 pipe_q2_f3 pipe_eg_q2_f3 eg_q2_f3_m4 pipe_q3_f0 pipe_q3_f3 pipe_eg_q3_f3 pipe_q1_f0; do
  timeout 40 ./bin/$b 0 5 >> $O 2>&1 || echo "$b failed ($?)" >> $O
done
cat $O
