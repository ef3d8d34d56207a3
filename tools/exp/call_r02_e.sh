# round 2, call e: fused ghost push v3 (per-direction flags, top chunk first / bottom chunk last, lazy wait) on 2 GPUs; TMA probe
python -m pytest tests/test_gpu_distributed.py -m gpu -x -q 2>&1 | tail -3
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for f in 1 0; do
  PMG_FUSED_HALO=$f $TR --master-port 2959$f bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu_fused$f.json 2> gpurun_out/bench_2gpu_fused$f.err
done
PMG_FUSED_ORDER=0 $TR --master-port 29593 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu_fused_natural.json 2> gpurun_out/bench_2gpu_fused_natural.err
python - <<'PY'
import json
for f in ("bench_2gpu_fused1","bench_2gpu_fused_natural","bench_2gpu_fused0"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, round(d["value"],3),"GDoF/s", round(d["ms_per_step"],3), "ms", d["per_level_ms"].get("applies_without_exchange"))
        for lv,row in list(zip(d["config"]["levels"], d["per_level_ms"]["ms"]))[-3:]: print("   ", lv, row)
    except Exception as e: print(f, "ERR", e)
PY
rm -f gpurun_out/exp_tma.txt; CUDA_VISIBLE_DEVICES=0 bash tools/exp/run_tma.sh 2>&1 | tail -30
