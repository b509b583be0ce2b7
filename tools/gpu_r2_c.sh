set -x
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused_graph_conv or gcn_stack" > gpurun_out/r2c_fused.log 2>&1; echo "fused rc=$?"; tail -40 gpurun_out/r2c_fused.log
nvidia-smi --query-gpu=name,memory.used --format=csv,noheader
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2c_tests.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/r2c_tests.log
