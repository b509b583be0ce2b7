set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2t_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2t_tests.log
timeout 600 python tools/gcn_layer_time.py > gpurun_out/r2t_layer_time.log 2>&1; tail -8 gpurun_out/r2t_layer_time.log
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err; echo "bench rc=$?"; cut -c1-260 gpurun_out/r2t_bench.json
