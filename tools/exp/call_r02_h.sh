# round 2, call h (2 GPUs): fused ghost push v4 (boundary planes copied after the march, not stored from the epilogue):
# parity, cost of the machinery alone on one GPU, bench on 2 GPUs
python -m pytest tests/test_gpu_distributed.py -m gpu -x -q 2>&1 | tail -3
CUDA_VISIBLE_DEVICES=0 SWEEP_DOFS=17e6 python tools/run_kernels.py 0 0 0 sweep > gpurun_out/selftest_plain.txt 2>&1
CUDA_VISIBLE_DEVICES=0 PMG_FUSED_SELFTEST=1 SWEEP_DOFS=17e6 python tools/run_kernels.py 0 0 0 sweep > gpurun_out/selftest_fused.txt 2>&1
paste -d'\n' gpurun_out/selftest_plain.txt gpurun_out/selftest_fused.txt | head -12 | sed 's/(0.* | / | /'
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$TR --master-port 29621 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu_fused_v4.json 2> gpurun_out/bench_2gpu_fused_v4.err
python - <<'PY'
import json
for f in ("bench_2gpu_fused_v4",):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, round(d["value"],3),"GDoF/s", round(d["ms_per_step"],3), "ms", d["per_level_ms"].get("applies_without_exchange"), "e2e", round(d["e2e"]["value"],3))
        for lv,row in list(zip(d["config"]["levels"], d["per_level_ms"]["ms"]))[-3:]: print("   ", lv, row)
    except Exception as e: print(f, "ERR", e)
PY
