set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_graph.py -m gpu -x -q > gpurun_out/tests_graph.log 2>&1; echo "graph tests rc=$?"; tail -25 gpurun_out/tests_graph.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_graph.json 2> gpurun_out/bench_graph.err; cat gpurun_out/bench_graph.json | cut -c1-250; tail -5 gpurun_out/bench_graph.err
