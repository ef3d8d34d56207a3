python -m pytest tests -m gpu -x -q 2>&1 | tail -5
bash tools/run_dist_drivers.sh 2 64 64
# fresh ncu captures of the shipped kernels (1 GPU), after the plain run
CUDA_VISIBLE_DEVICES=0 python tools/run_kernels.py 4 64 3 both > gpurun_out/rk_q4.log 2>&1 && \
CUDA_VISIBLE_DEVICES=0 ncu --set full --clock-control none --import-source on -k regex:pmg_plane -c 6 -o gpurun_out/prof_r02_final_plane_q4_c2 python tools/run_kernels.py 4 64 3 both > gpurun_out/ncu_q4.log 2>&1
tail -3 gpurun_out/ncu_q4.log
