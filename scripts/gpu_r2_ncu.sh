# round 2: one ncu --set full capture of the kernels matching $KREGEX (after the plain command exited 0)
set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
NUTSB_OVERLAP=0 $CMD > gpurun_out/plain.log 2>&1 &&
NUTSB_OVERLAP=0 ncu --set full --clock-control none --import-source on -k regex:"$KREGEX" -s ${SKIP:-1} -c ${COUNT:-2} -f -o gpurun_out/${OUT:-prof_r2} $CMD > gpurun_out/ncu_full.log 2>&1
tail -5 gpurun_out/ncu_full.log; ls -la gpurun_out/*.ncu-rep
