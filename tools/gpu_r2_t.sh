set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_nodal.py tests/test_gpu_relation.py tests/test_gpu_parity.py -m gpu -q -k "nodal or relation or bigru or single_stream" > gpurun_out/r2t_tests.log 2>&1; echo "tests rc=$?"; tail -30 gpurun_out/r2t_tests.log
