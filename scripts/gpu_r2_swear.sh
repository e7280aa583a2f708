# round 2: contains_swearing -- parity tests of both kernel forms, then the verdict kernel's time with and without k_ac_pair
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_swear_warp.py tests/test_speech.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest.log
for np in 0 1; do
  NUTSB_AC_NO_PAIR=$np timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_swear_$np.json 2> gpurun_out/r2_swear_$np.err; echo "bench rc=$?"
  python - <<PY
import json
d = json.loads(open('gpurun_out/r2_swear_$np.json').read().strip().splitlines()[-1])
print("NO_PAIR=$np ms_per_step", d["ms_per_step"], "verdict_kernel_ms", d["config"].get("verdict_kernel_ms"))
PY
done
