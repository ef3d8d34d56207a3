#!/bin/bash
# Ablation of the line-marching apply kernel: which of its parts costs what on its own (round 1: profiles/r01_v6_ablation_experiment.txt).
#   build (here, no GPU needed):   bash tools/exp/run_abl.sh build
#   run (on the GPU box):          gpurun --timeout 120 -- 'bash tools/exp/run_abl.sh'      -> gpurun_out/exp_abl.txt
# The switches are inserted into csrc/pmg_apply_sweep.h by tools/exp/ablation_hooks.py for the build only and removed again.
cd "$(dirname "$0")" || exit 1
if [ "$1" = build ]; then
  python ablation_hooks.py apply || exit 1
  trap 'python ablation_hooks.py revert' EXIT
  Q4="4 4 4 1 128 3 -DC_US=1 -DC_FM=0"
  ./build_exp.sh a_q4_base_f0 $Q4
  ./build_exp.sh q_q4_nostage $Q4 -DPMG_EXP_NOSTAGE
  ./build_exp.sh r_q4_nop12 $Q4 -DPMG_EXP_NOP12
  ./build_exp.sh s_q4_nop3 $Q4 -DPMG_EXP_NOP3
  ./build_exp.sh t_q4_nostage_nop12 $Q4 -DPMG_EXP_NOSTAGE -DPMG_EXP_NOP12
  ./build_exp.sh u_q4_nostage_nop3 $Q4 -DPMG_EXP_NOSTAGE -DPMG_EXP_NOP3
  ./build_exp.sh v_q4_nop12_nop3 $Q4 -DPMG_EXP_NOP12 -DPMG_EXP_NOP3
  ./build_exp.sh z_q4_us0_nop12_nop3 4 4 4 1 128 3 -DC_US=0 -DC_FM=0 -DPMG_EXP_NOP12 -DPMG_EXP_NOP3
  ./build_exp.sh skel_q4 $Q4 -DPMG_EXP_NOSTAGE -DPMG_EXP_NOP12 -DPMG_EXP_NOP3   # the skeleton alone: not measured in round 1
  ./build_exp.sh skel_q4_f3 4 4 4 1 128 3 -DC_US=1 -DC_FM=3 -DPMG_EXP_NOSTAGE -DPMG_EXP_NOP12 -DPMG_EXP_NOP3
  ./build_exp.sh w_q4_f3_nostage 4 4 4 1 128 3 -DC_US=1 -DC_FM=3 -DPMG_EXP_NOSTAGE
  ./build_exp.sh x_q2_nostage 2 8 8 2 128 3 -DC_US=1 -DC_FM=0 -DPMG_EXP_NOSTAGE
  ./build_exp.sh y_q2_nop12_nop3 2 8 8 2 128 3 -DC_US=1 -DC_FM=0 -DPMG_EXP_NOP12 -DPMG_EXP_NOP3
  exit 0
fi
cd bin || exit 1
mkdir -p ../../../gpurun_out
O=../../../gpurun_out/exp_abl.txt
for b in a_q4_base_f0 skel_q4 skel_q4_f3 q_q4_nostage r_q4_nop12 s_q4_nop3 t_q4_nostage_nop12 u_q4_nostage_nop3 v_q4_nop12_nop3 z_q4_us0_nop12_nop3 w_q4_f3_nostage x_q2_nostage y_q2_nop12_nop3; do
  [ -x ./$b ] || { echo "$b not built (bash tools/exp/run_abl.sh build)" >> $O; continue; }
  timeout 20 ./$b 0 5 >> $O 2>&1 || echo "$b failed" >> $O
done
cat $O
