set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/tests.log
timeout 600 python tools/step_profile.py > gpurun_out/step_profile.log 2>&1; tail -32 gpurun_out/step_profile.log
rm -f gpurun_out/step_trace.json
