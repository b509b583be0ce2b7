set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2t_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2t_tests.log
for i in 1 2 3 4 5 6; do timeout 300 python -m pytest tests/test_gpu_graph.py -m gpu -x -q 2>&1 | tail -1; done
