#!/bin/bash
# round 2, call C: Rosenbrock gradient kernels with the value factorisation (parity + A/B), copy ceiling
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_r2c.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_r2c.log
V=dynamicoed.jl_b200/_variants
rm -f $O/stiff_r2c.jsonl
for lib in $V/libdynoed_old.so $V/libdynoed_ndir1.so ""; do
  DYNOED_LIB=$lib timeout 600 python scripts/stiff_probe.py 2>&1 | grep '^{' | tee -a $O/stiff_r2c.jsonl
done
timeout 300 python scripts/pcie_probe.py > $O/pcie_1gpu.json 2> $O/pcie_1gpu.err; cat $O/pcie_1gpu.json
nvidia-smi topo -m > $O/topo.txt 2>&1; lscpu | head -30 > $O/lscpu.txt
