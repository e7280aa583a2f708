# round 2: everything the round-end driver runs, plus the profiles (each ncu run only after the plain command exited 0)
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest_gpu rc=$?"; tail -3 gpurun_out/r2_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"; cat gpurun_out/r2_smoke.log | tail -2
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; echo "bench_ref rc=$?"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc=$?"
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
NUTSB_OVERLAP=0 $CMD > gpurun_out/plain.log 2>&1 &&
NUTSB_OVERLAP=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
NUTSB_OVERLAP=0 $CMD > gpurun_out/plain2.log 2>&1 &&
NUTSB_OVERLAP=0 ncu --set full --clock-control none --import-source on -k regex:"k_fanout|k_render|k_direct|k_measure|k_ac_pair|k_ac_match|k_set_match|k_plan|k_rs_scatter|k_entry_info" -s 12 -c 14 -f -o gpurun_out/prof_r2_main $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
python scripts/show_list.py gpurun_out/r2_launches.csv | sort -rn | head -12
