# round 2, call b: fused ghost push on 2 GPUs -- parity, then bench with the push on / off, driver profile, 1-GPU regression check
python -m pytest tests/test_gpu_distributed.py -m gpu -x -q 2>&1 | tail -5
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for f in 1 0; do
  PMG_FUSED_HALO=$f $TR --master-port 2955$f bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu_fused$f.json 2> gpurun_out/bench_2gpu_fused$f.err
  tail -c 600 gpurun_out/bench_2gpu_fused$f.err
done
timeout 600 $TR --master-port 29561 --no-python portable-multigrid_b200/bin/polynomial_multigrid --dim 3 --hp 1 --degree 4 --profile 1 --cells 64 > gpurun_out/driver_c2_2gpu_fused.txt 2>&1
tail -12 gpurun_out/driver_c2_2gpu_fused.txt
CUDA_VISIBLE_DEVICES=0 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_1gpu_fusedbuild.json 2> gpurun_out/bench_1gpu_fusedbuild.err
