set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_gcnii.py -m gpu -q > gpurun_out/r2t_tests.log 2>&1; echo "tests rc=$?"; tail -40 gpurun_out/r2t_tests.log
