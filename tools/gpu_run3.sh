set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "aggregate or spmm or long_dialogues" > gpurun_out/tests_agg.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/tests_agg.log
timeout 300 python tools/spmm_variant.py > gpurun_out/spmm_variant.log 2>&1; cat gpurun_out/spmm_variant.log
timeout 300 python tools/spmm_phases.py > gpurun_out/spmm_phases.log 2>&1; tail -42 gpurun_out/spmm_phases.log
