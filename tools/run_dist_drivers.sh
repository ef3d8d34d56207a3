#!/bin/bash
# Distributed runs of the C drivers and the multi-GPU parity check on N GPUs of one box (N = $1, default 2):
#   gpurun --gpus N -- 'bash tools/run_dist_drivers.sh N [CELLS_C5] [CELLS_C2] [CELLS_C4]'
# Writes gpurun_out/dist_check_${N}gpu.log, driver_c5_${N}gpu.txt, driver_c2_${N}gpu.txt (per-level breakdown incl. halo).
N=${1:-2}; C5=${2:-64}; C2=${3:-64}; C4=${4:-0}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
BIN=portable-multigrid_b200/bin
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,P2P $TR --master-port 29541 tools/dist_check.py > gpurun_out/dist_check_${N}gpu.log 2>&1
echo "dist_check rc=$?"
timeout 600 $TR --master-port 29542 --no-python $BIN/polynomial_multigrid --dim 3 --hp 1 --degree 5 --coefficient 1 --tol 1e-10 --profile 1 --cells $C5 > gpurun_out/driver_c5_${N}gpu.txt 2>&1
echo "driver c5 rc=$?"; tail -25 gpurun_out/driver_c5_${N}gpu.txt
timeout 600 $TR --master-port 29543 --no-python $BIN/polynomial_multigrid --dim 3 --hp 1 --degree 4 --profile 1 --cells $C2 > gpurun_out/driver_c2_${N}gpu.txt 2>&1
echo "driver c2 rc=$?"; tail -22 gpurun_out/driver_c2_${N}gpu.txt
if [ "$C4" -gt 0 ]; then
  # BASELINE config 4: Q3, geometric multigrid, 160^3 cells per GPU (N = 8: 320^3 cells, 887 M DoFs)
  timeout 600 $TR --master-port 29544 --no-python $BIN/geometric_multigrid --degree 3 --profile 1 --cells $C4 > gpurun_out/driver_c4_${N}gpu.txt 2>&1
  echo "driver c4 rc=$?"; tail -22 gpurun_out/driver_c4_${N}gpu.txt
fi
