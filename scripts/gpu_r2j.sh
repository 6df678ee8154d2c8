#!/bin/bash
# round 2, 8-GPU call (tag = $1): copy ceiling (plain / NUMA-bound) and bench.py at N = 8 and N = 2
TAG=${1:-r2j}
O=gpurun_out; mkdir -p $O
nvidia-smi topo -m > $O/topo8.txt 2>&1; lscpu | head -25 > $O/lscpu8.txt; nproc >> $O/lscpu8.txt
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
    scripts/pcie_probe.py > $O/pcie_8gpu.json 2> $O/pcie_8gpu.err; echo "pcie N=8 rc=$?"; cut -c1-600 $O/pcie_8gpu.json
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 8 --steps 100 --warmup 5 > $O/bench_${TAG}_n8.json 2> $O/bench_${TAG}_n8.err; echo "bench N=8 rc=$?"; tail -3 $O/bench_${TAG}_n8.err | cut -c1-300
CUDA_VISIBLE_DEVICES=0,1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 \
    bench.py --gpus 2 --steps 100 --warmup 5 > $O/bench_${TAG}_n2.json 2> $O/bench_${TAG}_n2.err; echo "bench N=2 rc=$?"
for N in 2 8; do python - <<PY
import json
try:
    d = json.loads(open("$O/bench_${TAG}_n$N.json").read().strip().splitlines()[-1])
    print($N, "value %.4g ms %.4f e2e %.4g ceil %.4g obj-only %.4g multistart_ms %.3f verified %s %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["copy_ceiling"], d["e2e"]["value_objective_only"], d["multistart_1M_ms"], d["argmin_verified"], d["multistart_winner_verified"]))
except Exception as e:
    print($N, "no line:", e)
PY
done
