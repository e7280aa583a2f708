# A/B runs of kernel variants: rebuild on the GPU box with different -D flags, short bench each.
# usage: VARIANTS="-DX=1|-DX=2 -DY=3" bash scripts/gpu_matrix.sh
mkdir -p gpurun_out; : > gpurun_out/matrix.txt
IFS='|' read -ra VS <<< "${VARIANTS:--DNUTSB_TILE_OPS=128}"
for v in "${VS[@]}"; do
  NUTSB_NVCC_EXTRA="$v" python nuts333_b200/build.py --force > /dev/null 2>gpurun_out/build.err || { echo "$v BUILD FAILED" >> gpurun_out/matrix.txt; continue; }
  env $MENV timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>gpurun_out/m.err | python scripts/bench_line.py "$v" >> gpurun_out/matrix.txt 2>&1
done
cat gpurun_out/matrix.txt
