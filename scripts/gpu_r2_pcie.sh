# round 2: PCIe ceilings with N GPUs copying at once + the bench's host-buffer legs on N GPUs
N=${N:-8}
set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo_${N}gpu.txt 2>&1; (numactl -H || lscpu | grep -i numa) >> gpurun_out/r2_topo_${N}gpu.txt 2>&1; nproc >> gpurun_out/r2_topo_${N}gpu.txt
for n in 1 2 4 8; do
  if [ $n -le $N ]; then
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 scripts/probes/pcie_probe_multi.py 2>/dev/null | tail -n 1 > gpurun_out/r2_pcie_probe_${n}gpu.json; echo "probe $n rc=$?"; cat gpurun_out/r2_pcie_probe_${n}gpu.json
  fi
done
