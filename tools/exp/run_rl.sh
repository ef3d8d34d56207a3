#!/bin/bash
# one-shot timing of the rolled-loop / tile experiments: tools/exp/run_rl.sh  -> gpurun_out/exp_rl.txt
cd "$(dirname "$0")/bin" || exit 1
mkdir -p ../../../gpurun_out
for b in a_q4_base_f0 c_q4_rl_m4_f0 b_q4_rl_f0 o_q4_base_f3 h_q4_rl_f3 d_q4_l2_256_rl_f0 f_q2_rl_f0 g_q3_rl_f0 e_q4_l2_192_rl_f0 i_q2_rl_f3 j_q3_rl_f3 k_q3_l2_160_rl_f0 l_q2_l4_192_rl_f0 m_q4_l2_256_f0 n_q5_rl_f0 p_q1_rl_f0; do
  timeout 20 ./$b 0 5 >> ../../../gpurun_out/exp_rl.txt 2>&1 || echo "$b failed" >> ../../../gpurun_out/exp_rl.txt
done
cat ../../../gpurun_out/exp_rl.txt
