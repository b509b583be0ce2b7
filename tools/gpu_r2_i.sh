set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused or gcn_stack" > gpurun_out/r2i_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2i_tests.log
timeout 300 python tools/gcn_layer_phases.py > gpurun_out/r2i_phases.log 2>&1; echo "phases rc=$?"; grep -v "^   " gpurun_out/r2i_phases.log | tail -30
