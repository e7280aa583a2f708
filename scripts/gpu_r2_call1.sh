# round 2, call 1: full GPU parity suite (incl. the drop-in test) + a short bench to re-baseline on this box
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c1_pytest.log 2>&1; echo "pytest rc=$?" > gpurun_out/r2c1_rc.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2c1_bench.json 2> gpurun_out/r2c1_bench.err; echo "bench rc=$?" >> gpurun_out/r2c1_rc.txt
tail -5 gpurun_out/r2c1_pytest.log; cat gpurun_out/r2c1_bench.json; tail -3 gpurun_out/r2c1_bench.err; cat gpurun_out/r2c1_rc.txt
