set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
free -g | head -2 >> gpurun_out/gpu.txt; nproc >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q -k "not full_size" > gpurun_out/pytest_small.log 2>&1; echo "pytest_small rc=$?" >> gpurun_out/rc.txt
timeout 600 python bench.py --steps 3 --warmup 2 > gpurun_out/bench1.json 2> gpurun_out/bench1.err; echo "bench rc=$?" >> gpurun_out/rc.txt
timeout 1200 python -m pytest tests -m gpu -x -q -k "full_size" > gpurun_out/pytest_full.log 2>&1; echo "pytest_full rc=$?" >> gpurun_out/rc.txt
tail -5 gpurun_out/pytest_small.log; cat gpurun_out/bench1.json; tail -5 gpurun_out/bench1.err; tail -5 gpurun_out/pytest_full.log; cat gpurun_out/rc.txt
