#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 36 python -m pytest tests -m gpu -x -q -k "random_designs_vs_cpp_oracle and lv_" > $O/pytest_last2.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_last2.log
