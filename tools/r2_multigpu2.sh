#!/bin/bash
# final build on the 8-GPU box: the driver's protocol (--steps 20 --warmup 5) and a long run, N = 1 and 8
mkdir -p gpurun_out
OUT=gpurun_out/r2_scale_final.jsonl
: > $OUT
run() { local n=$1; shift
  if [ "$n" = 1 ]; then python bench.py --gpus 1 "$@"
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $n "$@"; fi
}
for n in 1 8; do
  run $n --steps 20 --warmup 5 --no-cpu 2>/dev/null | grep '^{' >> $OUT
  run $n --steps 20 --warmup 5 --no-cpu --no-e2e 2>/dev/null | grep '^{' >> $OUT
  run $n --steps 200 --warmup 20 --no-cpu --no-e2e 2>/dev/null | grep '^{' >> $OUT
done
run 8 --workloads sym_solve6,sym_invert6,sym_solve10 --steps 20 --warmup 5 --no-cpu --no-e2e 2>/dev/null | grep '^{' >> $OUT
run 4 --steps 20 --warmup 5 --no-cpu --no-e2e 2>/dev/null | grep '^{' >> $OUT
run 2 --steps 20 --warmup 5 --no-cpu --no-e2e 2>/dev/null | grep '^{' >> $OUT
wc -l $OUT
