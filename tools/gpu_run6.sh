set -x
mkdir -p gpurun_out
timeout 300 python tools/gru_time.py > gpurun_out/gru_time.log 2>&1; cat gpurun_out/gru_time.log
timeout 300 python -m pytest tests -m gpu -x -q -k "gru or GRU or model or golden" > gpurun_out/tests_gru.log 2>&1; tail -3 gpurun_out/tests_gru.log
