# round 2, final 1-GPU call: smoke, bench (both arms), launch list of the bench under ncu (shares per kernel), config 1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/bench_1gpu_final.json 2> gpurun_out/bench_1gpu_final.err; tail -c 300 gpurun_out/bench_1gpu_final.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_1gpu_final_reference.json 2> gpurun_out/bench_1gpu_final_reference.err
python bench.py --config c1 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_1gpu_final_c1.json 2> gpurun_out/bench_1gpu_final_c1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 1266 --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
python tools/launch_summary.py gpurun_out/r02_bench_launches.csv 211 | head -24
python - <<'PY'
import json
for f in ("bench_1gpu_final","bench_1gpu_final_reference","bench_1gpu_final_c1"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, {k:(round(v,4) if isinstance(v,float) else v) for k,v in d.items() if k in ("value","ms_per_step","apply_gdofs","apply_hbm_frac","gpu_launches","impl")}, "e2e", d["e2e"]["value"], "roofline", d.get("roofline",{}).get("frac"), d.get("roofline",{}).get("traffic"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e: print(f, "ERR", e)
PY
