set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gemm" > gpurun_out/r2o_tests.log 2>&1; echo "gemm tests rc=$?"; tail -5 gpurun_out/r2o_tests.log
timeout 400 python tools/umma_check.py > gpurun_out/r2o_umma_check.log 2>&1; echo "check rc=$?"; head -36 gpurun_out/r2o_umma_check.log
