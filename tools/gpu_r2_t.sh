set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2t_tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r2t_tests.log
