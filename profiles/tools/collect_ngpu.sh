#!/bin/bash
# Multi-GPU bench lines (through `gpurun --gpus N`): the c2 train step, launched the way the driver launches it, and the c4
# eval-only config.  Output: gpurun_out/<tag>_bench_<N>gpu.json, <tag>_bench_c4_<N>gpu.json
N=${1:-2}
TAG=${2:-r02}
O=gpurun_out
mkdir -p $O
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@"; }
run --steps 20 --warmup 5 > $O/${TAG}_bench_${N}gpu.json 2> $O/${TAG}_bench_${N}gpu.err
run --config c4 --steps 20 --warmup 5 > $O/${TAG}_bench_c4_${N}gpu.json 2> $O/${TAG}_bench_c4_${N}gpu.err
python - <<PY
import json
for f in ("$O/${TAG}_bench_${N}gpu.json", "$O/${TAG}_bench_c4_${N}gpu.json"):
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l); print(f, round(d["value"]), d["ms_per_step"], round(d["e2e"]["value"]))
PY
