set -e
for v in "" "-DNUTSB_DIR_MINBLOCKS=4 -DNUTSB_FD_MINBLOCKS=4" "-DNUTSB_DIR_MINBLOCKS=6 -DNUTSB_FD_MINBLOCKS=6" "-DNUTSB_FAN_MINBLOCKS=4 -DNUTSB_FD_MINBLOCKS=4 -DNUTSB_DIR_MINBLOCKS=4"; do
  NUTSB_NVCC_EXTRA="$v" python -m nuts333_b200.build --force > /dev/null
  python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('VARIANT [$v]', round(d['ms_per_step'],4), {k: round(x,3) for k,x in d['config']['kernel_ms_alone'].items() if isinstance(x,float)}, d['config']['timed_region_ms']['fanout_and_direct'])"
done
