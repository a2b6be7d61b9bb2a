#!/bin/bash
# N = 2 lines of the CIFAR shapes on the final build (the earlier N = 2 run predates the generic-GEMM changes; the graph-captured
# (2,6),(2,24) step hung then because only rank 0 ran the eager launch-count step)
cd "$GRAFT_REPO_ROOT"
N=2
mkdir -p gpurun_out
run() { name=$1; shift; timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/mg_${name}_n$N.json 2> gpurun_out/mg_${name}_n$N.err; python - <<Q
import json
try:
    d=json.load(open("gpurun_out/mg_${name}_n$N.json"))
    print("$name N=$N", round(d["value"]), d["unit"], round(d["ms_per_step"],3), "ms", "scaling", d.get("scaling"), "e2e", round(d["e2e"]["value"]))
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/mg_${name}_n$N.err").read()[-800:])
Q
}
run cifar26_weak --workload cifar_2_6__2_24
run cifar212_weak --workload cifar_2_12__2_24
run cifar223_weak --workload cifar_2_23__2_24
