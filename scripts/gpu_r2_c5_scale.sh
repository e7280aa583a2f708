# round 2: config 5 at N GPUs (strong scaling, chunks of 1M messages per GPU) and the default bench at N GPUs
N=${N:-8}
set -x
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --config c5 --c5-chunk 1000000 --gpus $N --steps 2 --warmup 3 2> gpurun_out/r2_bench_c5_${N}gpu.err | tail -n 1 > gpurun_out/r2_bench_c5_${N}gpu.json; echo "c5 rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_c5_${N}gpu.json')); print(d['n_gpus'], d['value'], d['ms_per_step'], d['config']['digest_checksum'], d['config']['deliveries_per_job'])"
if [ -n "$DEFAULT" ]; then
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --steps 5 --warmup 3 2> gpurun_out/r2_bench_${N}gpu.err | tail -n 1 > gpurun_out/r2_bench_${N}gpu.json; echo "default rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_${N}gpu.json')); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e_iov']['value'], d['e2e_speech_iov']['value'], d['e2e_speech_iov']['library_pipe'])"
fi
