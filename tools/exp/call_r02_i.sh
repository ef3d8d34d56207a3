# call i (1 GPU): fused machinery alone, natural chunk order (the default) vs top-first
SWEEP_DOFS=17e6 python tools/run_kernels.py 0 0 0 sweep > gpurun_out/selftest_plain.txt 2>&1
PMG_FUSED_SELFTEST=1 SWEEP_DOFS=17e6 python tools/run_kernels.py 0 0 0 sweep > gpurun_out/selftest_fused_natural.txt 2>&1
PMG_FUSED_ORDER=1 PMG_FUSED_SELFTEST=1 SWEEP_DOFS=17e6 python tools/run_kernels.py 0 0 0 sweep > gpurun_out/selftest_fused_topfirst.txt 2>&1
paste -d'\n' gpurun_out/selftest_plain.txt gpurun_out/selftest_fused_natural.txt gpurun_out/selftest_fused_topfirst.txt | head -18 | sed 's/(0.* | / | /'
