# launch list + full captures of the main kernels (each ncu run only after the plain command exited 0).
# Serial schedule (NUTSB_OVERLAP=0) so that the per-kernel figures are each kernel's own.
set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
NUTSB_OVERLAP=0 $CMD > gpurun_out/plain.log 2>&1 &&
NUTSB_OVERLAP=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
NUTSB_OVERLAP=0 $CMD > gpurun_out/plain2.log 2>&1 &&
NUTSB_OVERLAP=0 ncu --set full --clock-control none --import-source on -k regex:"k_fanout|k_render|k_direct|k_measure|k_ac_match|k_plan" -c 14 -f -o gpurun_out/prof_main $CMD > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | head -40
# the gather-list path (nutsb_write_batch_iov / nutsb_speech_batch_iov): launch list of the phase probe
python scripts/probes/e2e_phases.py > gpurun_out/e2e_phases.json 2> gpurun_out/e2e_phases.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_iov.csv python scripts/probes/e2e_phases.py > gpurun_out/ncu_list_iov.log 2>&1
