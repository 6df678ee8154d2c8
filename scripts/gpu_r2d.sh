#!/bin/bash
# round 2, call D (8 GPUs): copy ceiling and bench.py at N = 2, 4, 8 on one box
O=gpurun_out; mkdir -p $O
nvidia-smi topo -m > $O/topo8.txt 2>&1; lscpu | head -25 > $O/lscpu8.txt; nproc >> $O/lscpu8.txt
P=29500
for N in 2 4 8; do
  P=$((P+1))
  DEV=$(seq -s, 0 $((N-1)))
  CUDA_VISIBLE_DEVICES=$DEV timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P \
      scripts/pcie_probe.py > $O/pcie_${N}gpu.json 2> $O/pcie_${N}gpu.err; echo "pcie N=$N rc=$?"; cut -c1-400 $O/pcie_${N}gpu.json
  P=$((P+1))
  CUDA_VISIBLE_DEVICES=$DEV timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P \
      bench.py --gpus $N --steps 100 --warmup 5 > $O/bench_r2d_n$N.json 2> $O/bench_r2d_n$N.err; echo "bench N=$N rc=$?"; tail -3 $O/bench_r2d_n$N.err | cut -c1-300
done
timeout 240 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > $O/bench_r2d_n1.json 2> $O/bench_r2d_n1.err; echo "bench N=1 rc=$?"
for N in 1 2 4 8; do python - <<PY
import json
try:
    d = json.loads(open("$O/bench_r2d_n$N.json").read().strip().splitlines()[-1])
    print($N, "value %.4g ms %.4f e2e %.4g ceil %.4g obj-only %.4g multistart_ms %.3f verified %s %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["copy_ceiling"], d["e2e"]["value_objective_only"], d["multistart_1M_ms"], d["argmin_verified"], d["multistart_winner_verified"]))
except Exception as e:
    print($N, "no line:", e)
PY
done
