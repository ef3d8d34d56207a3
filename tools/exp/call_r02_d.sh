# round 2, call d: the whole box (8 GPUs) -- distributed parity, configs 5 / 2 / 4 through the C drivers with the per-level
# breakdown, bench.py (config 2 weak-scaled) with the fused ghost push on and off
N=${1:-8}
bash tools/run_dist_drivers.sh $N ${2:-160} ${3:-128} ${4:-320}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
PMG_TILE_VARIANT=6 $TR --master-port 29545 tools/dist_check.py 2>&1 | grep -v "NCCL INFO" > gpurun_out/dist_check_${N}gpu_plane_fused.log; tail -3 gpurun_out/dist_check_${N}gpu_plane_fused.log
for f in 1 0; do
  PMG_FUSED_HALO=$f $TR --master-port 2958$f bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${N}gpu_fused$f.json 2> gpurun_out/bench_${N}gpu_fused$f.err
done
python - $N <<'PY'
import json,sys
N=sys.argv[1]
for f in ("bench_%sgpu_fused1"%N,"bench_%sgpu_fused0"%N):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, round(d["value"],3),"GDoF/s", round(d["ms_per_step"],3), "ms", "e2e", round(d["e2e"]["value"],3), "step", round(d["roofline"]["ms_per_launch"],4), d["per_level_ms"].get("applies_without_exchange"), "cg", d["cg_solve"]["iterations"], round(d["cg_solve"]["ms"],1))
        for lv,row in zip(d["config"]["levels"], d["per_level_ms"]["ms"]): print("   ", lv, row)
    except Exception as e: print(f, "ERR", e)
PY
