mkdir -p gpurun_out; : > gpurun_out/inflight.txt
for n in 1 2 3; do
  timeout 300 python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-e2e --in-flight $n 2>gpurun_out/m.err | python scripts/bench_line.py "in-flight=$n" >> gpurun_out/inflight.txt 2>&1
done
cat gpurun_out/inflight.txt; tail -3 gpurun_out/m.err
