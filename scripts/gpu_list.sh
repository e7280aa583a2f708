# launch list of one bench step (serial schedule: NUTSB_OVERLAP=0 so that per-kernel times are not mixed)
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
NUTSB_OVERLAP=0 $CMD > gpurun_out/plain.log 2>&1 &&
NUTSB_OVERLAP=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
