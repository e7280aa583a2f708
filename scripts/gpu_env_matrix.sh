# A/B runs of run-time knobs (environment variables read by nutsb_create): short bench each.
# usage: ENVS="NUTSB_RENDER_PAD=0|NUTSB_RENDER_PAD=38000 NUTSB_SIDE_RENDER=3" bash scripts/gpu_env_matrix.sh
mkdir -p gpurun_out; : > gpurun_out/env_matrix.txt
IFS='|' read -ra VS <<< "${ENVS:-NUTSB_OVERLAP=1}"
for v in "${VS[@]}"; do
  env $v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>gpurun_out/m.err | python scripts/bench_line.py "$v" >> gpurun_out/env_matrix.txt 2>&1
done
cat gpurun_out/env_matrix.txt
