# call j (1 GPU): which part of the ticket costs?  system fence / gpu fence / no fence
for b in 0 8 4; do
PMG_FUSED_SELFTEST_BITS=$b PMG_FUSED_SELFTEST=1 SWEEP_DOFS=17e6 python tools/run_kernels.py 0 0 0 sweep 2>&1 | head -4 | sed "s/(0.* | / | /; s/^/bits=$b /"
done
