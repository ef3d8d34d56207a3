# round 2, call f: the whole box again -- fused ghost push v3 (top chunk first, lazy wait, per-direction flags; small slabs on the
# plane kernel), natural chunk order for comparison; config 5 with the stand-alone exchange by size (push kernel for large ones)
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29601 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${N}gpu_fused_v3.json 2> gpurun_out/bench_${N}gpu_fused_v3.err
PMG_FUSED_ORDER=0 $TR --master-port 29602 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${N}gpu_fused_v3_natural.json 2> gpurun_out/bench_${N}gpu_fused_v3_natural.err
python - $N <<'PY'
import json,sys
N=sys.argv[1]
for f in ("bench_%sgpu_fused_v3"%N,"bench_%sgpu_fused_v3_natural"%N):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, round(d["value"],3),"GDoF/s", round(d["ms_per_step"],3), "ms", "e2e", round(d["e2e"]["value"],3), "step", round(d["roofline"]["ms_per_launch"],4), d["per_level_ms"].get("applies_without_exchange"), "cg", d["cg_solve"]["iterations"], round(d["cg_solve"]["ms"],1))
        for lv,row in list(zip(d["config"]["levels"], d["per_level_ms"]["ms"]))[-5:]: print("   ", lv, row)
    except Exception as e: print(f, "ERR", e)
PY
BIN=portable-multigrid_b200/bin
for mb in 0 1000000; do
  PMG_P2P_MIN_BYTES=$mb timeout 600 $TR --master-port 2961$((mb>0)) --no-python $BIN/polynomial_multigrid --dim 3 --hp 1 --degree 5 --coefficient 1 --tol 1e-10 --profile 1 --cells ${2:-160} > gpurun_out/driver_c5_${N}gpu_minbytes$mb.txt 2>&1
  echo "PMG_P2P_MIN_BYTES=$mb"; grep -A4 "solve time" gpurun_out/driver_c5_${N}gpu_minbytes$mb.txt
done
