# round 2: config 5 at N GPUs, the whole job in one chunk (one write batch per GPU)
N=${N:-8}
set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 bench.py --config c5 --c5-chunk 10000000 --gpus $N --steps 2 --warmup 3 2> gpurun_out/r2_c5_one_chunk_${N}gpu.err | tail -n 1 > gpurun_out/r2_c5_one_chunk_${N}gpu.json; echo "rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r2_c5_one_chunk_${N}gpu.json')); print(d['n_gpus'], d['value'], d['ms_per_step'], d['config']['digest_checksum'], d['gpu_launches'])"
