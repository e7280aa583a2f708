# full captures of selected kernels: KERNELS="k_render|k_direct" bash scripts/gpu_prof2.sh
set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"${KERNELS:-k_fanout}" -c ${COUNT:-4} -o gpurun_out/prof_sel -f $CMD > gpurun_out/ncu_sel.log 2>&1
ls -la gpurun_out | head -5
