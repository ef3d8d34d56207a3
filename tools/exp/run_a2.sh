#!/bin/bash
# one-shot timing of the A2 (phase-2 items alternate between the halves of the CTA) and RL + 4 CTAs/SM variants -> gpurun_out/exp_a2.txt
cd "$(dirname "$0")/bin" || exit 1
mkdir -p ../../../gpurun_out
O=../../../gpurun_out/exp_a2.txt
timeout 15 ./C_q4_a2_rl_m4_f0 0 5 2 >> $O 2>&1
for b in A_q4_a2_f0 B_q4_a2_f3 D_q2_a2_f0 E_q2_a2_f3 F_q3_a2_f0 G_q3_a2_f3 J_q2_rl_m4_f0 L_q2_a2_rl_m4_f0 M_q4_a2_rl_f3 H_q1_a2_f0 I_q1_a2_f3 K_q5_rl_m3_f0; do
  timeout 15 ./$b 0 5 >> $O 2>&1 || echo "$b failed" >> $O
done
cat $O
