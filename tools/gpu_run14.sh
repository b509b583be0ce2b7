set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_gated.py -m gpu -x -q > gpurun_out/tests_gated.log 2>&1; echo "gated rc=$?"; tail -25 gpurun_out/tests_gated.log
