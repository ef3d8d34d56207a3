# round 2, call g (1 GPU): what does the fused-exchange machinery cost by itself?  The same launches with the PUSH kernel
# instance, flag words and tickets but no neighbours (PMG_FUSED_SELFTEST=1) against the plain instance, ~17 M DoFs per degree
SWEEP_DOFS=17e6 python tools/run_kernels.py 0 0 0 sweep > gpurun_out/selftest_plain.txt 2>&1
PMG_FUSED_SELFTEST=1 SWEEP_DOFS=17e6 python tools/run_kernels.py 0 0 0 sweep > gpurun_out/selftest_fused.txt 2>&1
paste -d'\n' gpurun_out/selftest_plain.txt gpurun_out/selftest_fused.txt | head -14
