set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/r2t_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2t_tests.log
timeout 100 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/r2t_bench.json
