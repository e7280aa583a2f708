# gather-list entry point: its parity tests, the small suite, a bench with both host-buffer legs
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "not full_size" > gpurun_out/pytest_small.log 2>&1; echo "pytest_small rc=$?" > gpurun_out/rc.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/rc.txt
tail -4 gpurun_out/pytest_small.log; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err; cat gpurun_out/rc.txt
