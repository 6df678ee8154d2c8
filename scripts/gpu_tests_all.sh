O=gpurun_out; mkdir -p $O
timeout 1000 python -m pytest tests -m gpu -q > $O/pytest_r2tp2.log 2>&1; echo "pytest rc=$?"; tail -8 $O/pytest_r2tp2.log
