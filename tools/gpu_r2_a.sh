# round-2 call A: state check of HEAD on a fresh B200, variant-2 (any-length tcgen05 aggregate) validation,
# compute-sanitizer memcheck / racecheck over the tcgen05 + recurrence kernel tests.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2a_tests.log
timeout 240 python tools/spmm_variant.py > gpurun_out/r2a_spmm_variant.log 2>&1; echo "variant rc=$?"; tail -45 gpurun_out/r2a_spmm_variant.log
nvidia-smi --query-gpu=name,memory.used --format=csv
timeout 500 compute-sanitizer --tool memcheck --log-file gpurun_out/r2a_memcheck.log python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "aggregate or gemm_nn or bigru2 or spmm_and_grad or gcn_stack or adjacency" > gpurun_out/r2a_memcheck_pytest.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/r2a_memcheck_pytest.log; tail -5 gpurun_out/r2a_memcheck.log
timeout 400 compute-sanitizer --tool racecheck --log-file gpurun_out/r2a_racecheck.log python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "aggregate or gemm_nn or spmm_and_grad" > gpurun_out/r2a_racecheck_pytest.log 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/r2a_racecheck_pytest.log; tail -5 gpurun_out/r2a_racecheck.log
