#!/bin/bash
# the driver's launch line at N=8 (graph-timed region), then configs 3 and 5 in one more launch
mkdir -p gpurun_out
out=gpurun_out/r2_graph_8gpu.jsonl
: > $out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 >> $out 2> gpurun_out/r2_graph_8gpu.err
echo "rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 20 --warmup 5 --no-e2e --workloads sym_solve3,sym_solve6,sym_invert6,sym_solve10 >> $out 2>> gpurun_out/r2_graph_8gpu.err
echo "rc=$?"
python - <<'P'
import json
for l in open("gpurun_out/r2_graph_8gpu.jsonl"):
    if not l.startswith("{"): continue
    d = json.loads(l)
    print("%-14s n_gpus %d steps %3d  %-12s %8.2f us/step  %8.2f Gmat/s per-GPU frac %.3f  e2e %s" % (
        d["config"]["routine"] + str(d["config"]["n"]), d["n_gpus"], d["steps"], d["timing"]["launch"][:10],
        d["ms_per_step"] * 1e3, d["value"] / 1e9, d["roofline"]["frac"], d["e2e"] and "%.2f G" % (d["e2e"]["value"] / 1e9)))
P
grep -v "^$\|\*\*\*\|OMP_NUM" gpurun_out/r2_graph_8gpu.err | tail -5
