set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mfn.py -m gpu -x -q > gpurun_out/r2x_tests.log 2>&1; echo "tests rc=$?"; tail -30 gpurun_out/r2x_tests.log
