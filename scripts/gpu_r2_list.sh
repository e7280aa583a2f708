# round 2: targeted GPU tests ($TESTS), a short bench, and the launch list of one bench step (serial schedule)
set -x
mkdir -p gpurun_out
if [ -n "$TESTS" ]; then timeout 900 python -m pytest $TESTS -m gpu -x -q > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest.log; fi
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2_bench.json'))
print("ms_per_step", d["ms_per_step"], "value", d["value"], "kernel_ms_alone", d["config"].get("kernel_ms_alone"), "launches", d.get("gpu_launches"))
PY
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
NUTSB_OVERLAP=0 $CMD > gpurun_out/plain.log 2>&1 &&
NUTSB_OVERLAP=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
python scripts/show_list.py gpurun_out/launches.csv | sort -rn | head -40
