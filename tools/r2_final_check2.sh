#!/bin/bash
# final state of round 2: GPU tests, smoke, both bench arms (driver protocol), ncu launch list of the default command
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; tail -c 300 gpurun_out/r2_bench_default.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; tail -c 300 gpurun_out/r2_bench_reference.err
python bench.py --workload sym_solve3_1m --steps 20 --warmup 5 --no-cpu > gpurun_out/r2_bench_config1.json 2>> gpurun_out/r2_bench_default.err
B="python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_default_graph.csv $B > gpurun_out/r2_ncu_launches_graph.log 2>&1
echo "ncu rc=$?"
python profiles/summarize_ncu.py list gpurun_out/r2_launches_default_graph.csv > gpurun_out/r2_sym_solve3_launches_graph.txt 2>&1
tail -15 gpurun_out/r2_sym_solve3_launches_graph.txt
python - <<'P'
import json
for f in ("r2_bench_default", "r2_bench_reference", "r2_bench_config1"):
    d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
    print(f, "value %.4g" % d["value"], "ms/step %.5f" % d["ms_per_step"], "frac", d.get("roofline", {}).get("frac"), "e2e", d["e2e"]["value"] if d.get("e2e") else None,
          "launches", d.get("gpu_launches"), d.get("timing", {}).get("launch"), d.get("clocks"))
P
