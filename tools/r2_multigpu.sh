#!/bin/bash
# round 2, multi-GPU box: scaling of configs 2, 3, 5 at 1/2/4/8 GPUs (device-resident and e2e), the host's
# PCIe ceiling at 1/2/4/8 concurrent ranks, and the tests that need >= 2 devices.
mkdir -p gpurun_out
NG=${NG:-8}
OUT=gpurun_out/r2_scale
: > $OUT.jsonl
run() { # gpus, extra args...
  local n=$1; shift
  if [ "$n" = 1 ]; then python bench.py --gpus 1 "$@"
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n "$@"; fi
}
python -m pytest tests -m gpu -x -q -k "non_current_device or multi_gpu or host_pipeline" > gpurun_out/r2_pytest_multigpu.log 2>&1
echo "pytest rc=$? on $(nvidia-smi -L | wc -l) GPUs" >> gpurun_out/r2_pytest_multigpu.log
for n in 1 2 4 8; do
  [ $n -le $NG ] || continue
  if [ $n = 1 ]; then python tools/pcie_ceiling.py; else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 tools/pcie_ceiling.py; fi 2>/dev/null | grep '^{' >> gpurun_out/r2_pcie_ceiling.jsonl
done
for n in 1 2 4 8; do
  [ $n -le $NG ] || continue
  run $n --workload sym_solve3 --steps 200 --warmup 20 --no-cpu 2>/dev/null | grep '^{' >> $OUT.jsonl
  run $n --workloads sym_solve6,sym_invert6,sym_solve10,sym_solve3_1m --steps 100 --warmup 10 --no-cpu --no-e2e 2>/dev/null | grep '^{' >> $OUT.jsonl
done
# the driver's own protocol (20 steps) for config 2 at the largest N, and chunk-size sensitivity of the e2e path
run $NG --workload sym_solve3 --steps 20 --warmup 3 --no-cpu --no-e2e 2>/dev/null | grep '^{' >> $OUT.jsonl
for mb in 4 64; do
  NFM_HOST_CHUNK_MB=$mb run $NG --workload sym_solve3 --steps 20 --warmup 3 --no-cpu 2>/dev/null | grep '^{' | sed "s/^{/{\"chunk_mb\": $mb, /" >> gpurun_out/r2_e2e_chunk.jsonl
done
tail -3 gpurun_out/r2_pytest_multigpu.log; wc -l $OUT.jsonl gpurun_out/r2_pcie_ceiling.jsonl
