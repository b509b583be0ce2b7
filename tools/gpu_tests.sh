# GPU parity suite on a fresh B200 box; log under gpurun_out/ (first argument = tag)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/$1_tests.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/$1_tests.log
