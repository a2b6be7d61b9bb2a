#!/bin/bash
# N-GPU runs (N = $1): DP equivalence check, weak scaling of cfg2 and the CIFAR shapes, strong scaling of cfg2 (global 512)
cd "$GRAFT_REPO_ROOT"
N=${1:-2}
mkdir -p gpurun_out
run() { name=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/mg_${name}_n$N.json 2> gpurun_out/mg_${name}_n$N.err; python - <<Q
import json
try:
    d=json.load(open("gpurun_out/mg_${name}_n$N.json"))
    print("$name N=$N", round(d["value"]), d["unit"], round(d["ms_per_step"],3), "ms", "scaling", d.get("scaling"), "dp_check", d.get("dp_check"), "e2e", round(d["e2e"]["value"]))
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/mg_${name}_n$N.err").read()[-1500:])
Q
}
run cfg2_check --check-dp
run cfg2_weak
run cfg2_overlap --overlap
run cfg2_strong --strong --batch 512
run cifar26_weak --workload cifar_2_6__2_24
run cifar26_check --workload cifar_2_6__2_24 --check-dp --graph off
run cifar212_weak --workload cifar_2_12__2_24
run cifar223_weak --workload cifar_2_23__2_24
run cifar223_overlap --workload cifar_2_23__2_24 --overlap
run cifar223_strong --workload cifar_2_23__2_24 --strong --batch 64
run three_weak --workload three_eps
