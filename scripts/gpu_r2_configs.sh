# round 2: the bench lines of every config on one GPU
set -x
mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; echo "default rc=$?"
timeout 600 python bench.py --config c2 --steps 5 --warmup 3 > gpurun_out/r2_bench_c2.json 2> gpurun_out/r2_bench_c2.err; echo "c2 rc=$?"
timeout 600 python bench.py --config c4 --steps 5 --warmup 3 > gpurun_out/r2_bench_c4.json 2> gpurun_out/r2_bench_c4.err; echo "c4 rc=$?"
timeout 600 python bench.py --config c5 --c5-msgs 400000 --c5-users 10000 --steps 2 --warmup 3 > gpurun_out/r2_bench_c5_small.json 2> gpurun_out/r2_bench_c5_small.err; echo "c5 small rc=$?"
timeout 900 python bench.py --config c5 --steps 2 --warmup 3 > gpurun_out/r2_bench_c5_1gpu.json 2> gpurun_out/r2_bench_c5_1gpu.err; echo "c5 rc=$?"
tail -3 gpurun_out/r2_bench_*.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_bench_*.json')):
    try:
        d=json.load(open(f)); print(f, d["value"], d["unit"], d["ms_per_step"], d["config"].get("digest_checksum"), list(d.keys()))
        if "vs_cpu_components" in d: print(json.dumps(d["vs_cpu_components"], indent=1))
    except Exception as e: print(f, "ERR", e)
PY
