#!/bin/bash
# eager vs CUDA-graph launch of the K timed steps under the driver's protocol (--steps 20 --warmup 5),
# on one GPU with the slab sizes one rank sees at 8 / 4 / 2 / 1 GPUs of config 2
mkdir -p gpurun_out
out=gpurun_out/r2_graph_ab.jsonl
: > $out
for rep in 1 2 3; do
  for b in 2097152 4194304; do
    for mode in eager graph; do
      python bench.py --batch $b --steps 20 --warmup 5 --no-e2e --no-cpu --launch $mode >> $out 2>> gpurun_out/r2_graph_ab.err
    done
  done
done
for mode in eager graph; do
  python bench.py --batch 8388608 --steps 20 --warmup 5 --no-e2e --no-cpu --launch $mode >> $out 2>> gpurun_out/r2_graph_ab.err
  python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu --launch $mode >> $out 2>> gpurun_out/r2_graph_ab.err
  python bench.py --batch 2097152 --steps 200 --warmup 5 --no-e2e --no-cpu --launch $mode >> $out 2>> gpurun_out/r2_graph_ab.err
done
python - <<'P'
import json
for l in open("gpurun_out/r2_graph_ab.jsonl"):
    d = json.loads(l)
    print("%9d matrices  steps %3d  %-6s %8.2f us/step  frac %.3f  launches %d" % (
        d["config"]["batch"], d["steps"], d["timing"]["launch"].split(":")[0].split()[-1] if "graph" in d["timing"]["launch"] else "eager",
        d["ms_per_step"] * 1e3, d["roofline"]["frac"], d["gpu_launches"]))
P
tail -c 600 gpurun_out/r2_graph_ab.err
