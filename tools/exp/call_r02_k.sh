# call k (2 GPUs): the whole GPU suite with the final fused ghost push, then the 2-GPU bench
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$TR --master-port 29631 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_2gpu_final.json 2> gpurun_out/bench_2gpu_final.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_2gpu_final.json").read().strip().splitlines()[-1])
print(round(d["value"],3),"GDoF/s", round(d["ms_per_step"],3), "ms", d["per_level_ms"].get("applies_without_exchange"), "e2e", round(d["e2e"]["value"],3), "cg", d["cg_solve"]["iterations"], round(d["cg_solve"]["ms"],1))
for lv,row in list(zip(d["config"]["levels"], d["per_level_ms"]["ms"]))[-4:]: print("   ", lv, row)
PY
