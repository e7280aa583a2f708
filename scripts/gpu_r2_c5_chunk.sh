# round 2: config 5 on one GPU with larger chunks (fewer write batches per job)
set -x
mkdir -p gpurun_out
for ch in 2500000 5000000 10000000; do
  timeout 600 python bench.py --config c5 --c5-chunk $ch --steps 2 --warmup 2 2> gpurun_out/r2_c5_chunk_$ch.err | tail -n 1 > gpurun_out/r2_c5_chunk_$ch.json; echo "rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/r2_c5_chunk_$ch.json')); print($ch, d['value'], d['ms_per_step'], d['config']['digest_checksum'], d['gpu_launches'])" || tail -5 gpurun_out/r2_c5_chunk_$ch.err
done
