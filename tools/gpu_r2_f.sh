set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q -k "gemm or aggregate or fused or gcn_stack or bigru2 or golden" > gpurun_out/r2f_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2f_tests.log
timeout 300 python tools/gcn_layer_phases.py > gpurun_out/r2f_phases.log 2>&1; echo "phases rc=$?"; cat gpurun_out/r2f_phases.log | tail -20
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r2f_bench.json; tail -3 gpurun_out/r2f_bench.err
