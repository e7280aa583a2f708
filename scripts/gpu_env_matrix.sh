# A/B runs over environment settings (no rebuild): ENVS="A=1 B=2|A=2 B=2" bash scripts/gpu_env_matrix.sh
mkdir -p gpurun_out; : > gpurun_out/env_matrix.txt
IFS='|' read -ra VS <<< "${ENVS:-X=0}"
for v in "${VS[@]}"; do
  env $v timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>gpurun_out/m.err | python scripts/bench_line.py "$v" >> gpurun_out/env_matrix.txt 2>&1
done
cat gpurun_out/env_matrix.txt
