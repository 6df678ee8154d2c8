#!/bin/bash
# One gpurun call: GPU parity tests, bench line, ncu launch list + one full capture of the hot kernel.
# usage: scripts/gpu_round.sh <tag>
TAG=${1:-r1}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_$TAG.log
tail -3 $O/pytest_$TAG.log
python bench.py --steps 200 --warmup 5 > $O/bench_$TAG.log 2> $O/bench_$TAG.err; echo "bench rc=$?"
cat $O/bench_$TAG.log
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref_$TAG.log 2>&1; echo "ref rc=$?"
python scripts/occupancy_sweep.py > $O/occupancy_$TAG.log 2>&1; echo "occ rc=$?"
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$TAG.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/ncu_launches_$TAG.log 2>&1
echo "launch list rc=$?"
python scripts/prof_kernel.py 65536 > $O/plain_prof_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:oed_eval -s 3 -c 1 -f -o $O/prof_$TAG \
    python scripts/prof_kernel.py 65536 > $O/ncu_full_$TAG.log 2>&1
echo "full capture rc=$?"
cat $O/plain_prof_$TAG.log
