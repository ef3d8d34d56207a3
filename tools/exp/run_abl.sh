#!/bin/bash
# one-shot ablation of the line-marching kernel (hooks: tools/exp/ablation_hooks.patch): -> gpurun_out/exp_abl.txt
cd "$(dirname "$0")/bin" || exit 1
mkdir -p ../../../gpurun_out
O=../../../gpurun_out/exp_abl.txt
for b in q_q4_nostage r_q4_nop12 s_q4_nop3 t_q4_nostage_nop12 u_q4_nostage_nop3 v_q4_nop12_nop3 z_q4_us0_nop12_nop3 w_q4_f3_nostage x_q2_nostage y_q2_nop12_nop3; do
  timeout 20 ./$b 0 5 >> $O 2>&1 || echo "$b failed" >> $O
done
timeout 20 ./c_q4_rl_m4_f0 0 5 2 >> $O 2>&1
timeout 20 ./c_q4_rl_m4_f0 0 5 3 >> $O 2>&1
timeout 20 ./a_q4_base_f0 0 5 3 >> $O 2>&1
cat $O
