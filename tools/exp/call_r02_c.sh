# round 2, call c: DMMA cell-operator experiment (1 GPU); per-level breakdown of the bench cycle at 1 and 2 GPUs, fused push on / off
CUDA_VISIBLE_DEVICES=0 tools/exp/bin/exp_dmma > gpurun_out/exp_dmma.txt 2>&1; cat gpurun_out/exp_dmma.txt
CUDA_VISIBLE_DEVICES=0 tools/exp/bin/exp_dmma 262144 8 >> gpurun_out/exp_dmma.txt 2>&1; tail -4 gpurun_out/exp_dmma.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for f in 1 0; do
  PMG_FUSED_HALO=$f $TR --master-port 2957$f bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu_fused$f.json 2> gpurun_out/bench_2gpu_fused$f.err
done
CUDA_VISIBLE_DEVICES=0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_1gpu_levels.json 2> gpurun_out/bench_1gpu_levels.err
python - <<'PY'
import json
for f in ("bench_1gpu_levels","bench_2gpu_fused1","bench_2gpu_fused0"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"],3), "ms", d["per_level_ms"].get("applies_without_exchange"))
        for lv,row in zip(d["config"]["levels"], d["per_level_ms"]["ms"]): print("   ", lv, row)
    except Exception as e: print(f, "ERR", e)
PY
