# final scaling point of round 2 on N GPUs: bench.py as the driver runs it
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_${N}gpu_final.json 2> gpurun_out/bench_${N}gpu_final.err
python - $N <<'PY'
import json,sys
N=sys.argv[1]
d=json.loads(open("gpurun_out/bench_%sgpu_final.json"%N).read().strip().splitlines()[-1])
print(N, "GPUs:", round(d["value"],3),"GDoF/s", round(d["ms_per_step"],3), "ms", d["per_level_ms"].get("applies_without_exchange"), "e2e", round(d["e2e"]["value"],3), "cg", d["cg_solve"]["iterations"], round(d["cg_solve"]["ms"],1), d["clocks"])
for lv,row in list(zip(d["config"]["levels"], d["per_level_ms"]["ms"]))[-5:]: print("   ", lv, row)
PY
