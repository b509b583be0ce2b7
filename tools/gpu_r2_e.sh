set -x
mkdir -p gpurun_out
timeout 300 python tools/gcn_layer_phases.py > gpurun_out/r2e_phases.log 2>&1; echo "phases rc=$?"; cat gpurun_out/r2e_phases.log | tail -20
