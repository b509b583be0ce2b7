set -x
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused_graph_conv_layer" > gpurun_out/r2s_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2s_tests.log
timeout 240 python tools/gcn_layer_time.py 0 > gpurun_out/r2s_time.log 2>&1; echo "time rc=$?"; cat gpurun_out/r2s_time.log
timeout 240 python tools/gcn_layer2_phases.py > gpurun_out/r2s_phases.log 2>&1; echo "rc=$?"; head -12 gpurun_out/r2s_phases.log
