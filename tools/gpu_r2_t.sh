set -x
mkdir -p gpurun_out
timeout 240 python tools/gcn_layer2_phases.py > gpurun_out/r2t_phases.log 2>&1; echo "rc=$?"; cat gpurun_out/r2t_phases.log
