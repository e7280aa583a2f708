# round 2: config 5 (strong scaling) on N GPUs + the sharded parity test on real devices
N=${N:-2}
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multi.py -m gpu -x -q > gpurun_out/r2_multi_${N}gpu_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_multi_${N}gpu_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --config c5 --gpus $N --steps 2 --warmup 3 > gpurun_out/r2_bench_c5_${N}gpu.json 2> gpurun_out/r2_bench_c5_${N}gpu.err; echo "c5 rc=$?"
tail -n 3 gpurun_out/r2_bench_c5_${N}gpu.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2_bench_c5_${N}gpu.json')); print(d["n_gpus"], d["value"], d["ms_per_step"], d["config"]["digest_checksum"], d["config"]["deliveries_per_job"])
PY
