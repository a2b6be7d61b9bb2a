#!/bin/bash
# first GPU pass of the round: parity tests, smoke, bench lines, microbenchmark grid
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
tail -5 gpurun_out/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; echo "bench rc=$?"
for wl in cfg1 one_eps cifar_2_6__2_24 cifar_2_12__2_24 cifar_2_23__2_24 three_eps three_eps_32; do
  timeout 600 python bench.py --steps 20 --warmup 5 --workload $wl > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; echo "bench $wl rc=$?"
done
timeout 300 python bench.py --steps 20 --warmup 5 --workload cfg1 --graph off --no-cpu-baseline > gpurun_out/bench_cfg1_nograph.json 2> gpurun_out/bench_cfg1_nograph.err
timeout 300 python bench.py --steps 20 --warmup 5 --workload cifar_2_6__2_24 --graph off --no-cpu-baseline > gpurun_out/bench_cifar_2_6__2_24_nograph.json 2> gpurun_out/bench_cifar26_nograph.err
timeout 900 python benchmarks/eps_microbench.py --json gpurun_out/microbench_grid.json > gpurun_out/microbench_grid.log 2>&1; echo "microbench rc=$?"
tail -50 gpurun_out/microbench_grid.log
