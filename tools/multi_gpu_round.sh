#!/bin/bash
# Multi-GPU session on N GPUs of one box: strong-scaling bench (default), other modes, training step, density grid.
N=${1:-8}; TAG=${TAG:-r02}
mkdir -p gpurun_out
run() { name=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 "$@" > gpurun_out/${TAG}_${name}_${N}gpu.json 2> gpurun_out/${TAG}_${name}_${N}gpu.err; echo "$name exit $?"; }
run bench bench.py --gpus $N --steps 10 --warmup 3
run bench_thuman_strong bench.py --gpus $N --steps 10 --warmup 3 --workload thuman --mode strong
if [ "${FULL:-1}" = "1" ]; then
run bench_thuman_weak bench.py --gpus $N --steps 10 --warmup 3 --workload thuman --mode weak
run train bench.py --gpus $N --steps 10 --warmup 3 --workload train
run grid256 tools/grid_query.py --res 256
fi
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${TAG}_*_${N}gpu.json")):
    for l in open(f):
        try: d=json.loads(l)
        except Exception: continue
        if "ms_per_step" in d:
            print(f.split("/")[-1], d["config"]["workload"], d["scaling"], "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],3), "value", round(d["value"]), d.get("single_gpu_same_workload",{}).get("ms_per_step"), d.get("active_points_per_gpu"))
        else:
            print(f.split("/")[-1], d)
PY
